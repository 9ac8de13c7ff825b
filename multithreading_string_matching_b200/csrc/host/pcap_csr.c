/* pcap_csr.c -- savefile ingest: the step before the hot path (serial.c:91-141, openmp_data.c:94-147).
 *
 * The reference lets libpcap frame the records (pcap_open_offline serial.c:91, pcap_next_ex :115),
 * copies every frame, extracts its payload and stores one malloc'd buffer per payload.  Here the
 * savefile is mapped, walked twice (size, then copy) and the accepted payloads are packed back to
 * back into ONE flat buffer plus an offsets array -- the CSR batch the device consumes -- in pinned
 * memory so the H2D copies can run asynchronously.
 *
 * Frames are read with their captured length (openmp_data.c:114-116; serial.c:117-120 uses the wire
 * length, the same number whenever caplen == len).  Like the reference's `while (pcap_next_ex(...)
 * >= 0)` loop, a truncated trailing record ends the walk without an error.
 */
#include <errno.h>
#include <fcntl.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include "kmpb_internal.h"

#define PCAP_GLOBAL_HDR 24u
#define PCAP_RECORD_HDR 16u
#define CSR_TAIL_PAD 64u

typedef int (*extract_fn)(const uint8_t *, uint32_t, uint32_t *, uint32_t *);

static uint32_t load32(const uint8_t *p, int swapped)
{
    uint32_t v;
    memcpy(&v, p, sizeof v);
    return swapped ? __builtin_bswap32(v) : v;
}

/* One pass over the records.  With dst == NULL it only sizes the batch. */
static void walk(const uint8_t *file, size_t size, int swapped, extract_fn extract,
                 uint8_t *dst, uint64_t *offsets, uint64_t *n_packets, uint64_t *n_frames, uint64_t *total)
{
    size_t at = PCAP_GLOBAL_HDR;
    uint64_t packets = 0, frames = 0, bytes = 0;
    while (size - at >= PCAP_RECORD_HDR) {
        uint32_t caplen = load32(file + at + 8, swapped);
        at += PCAP_RECORD_HDR;
        if (caplen > size - at) break; /* truncated record: libpcap reports an error, the loop ends */
        uint32_t off, len;
        frames++;
        if (extract(file + at, caplen, &off, &len)) {
            if (dst) {
                memcpy(dst + bytes, file + at + off, len);
                offsets[packets] = bytes;
            }
            bytes += len;
            packets++;
        }
        at += caplen;
    }
    if (dst) offsets[packets] = bytes;
    *n_packets = packets;
    *n_frames = frames;
    *total = bytes;
}

int kmpb_load_pcap_csr(const char *path, int proto, int pinned, kmpb_csr *out)
{
    if (path == NULL || out == NULL) return kmpb_fail(KMPB_EINVAL, "kmpb_load_pcap_csr: NULL argument");
    if (proto != KMPB_PROTO_UDP && proto != KMPB_PROTO_TCP) return kmpb_fail(KMPB_EINVAL, "unknown protocol %d", proto);
    memset(out, 0, sizeof *out);
    int fd = open(path, O_RDONLY);
    struct stat st;
    if (fd < 0 || fstat(fd, &st) != 0) {
        int e = errno;
        if (fd >= 0) close(fd);
        kmpb_fail(KMPB_EIO, "%s: %s", path, strerror(e));
        errno = e;
        return KMPB_EIO;
    }
    size_t size = (size_t)st.st_size;
    if (size < PCAP_GLOBAL_HDR) {
        close(fd);
        return kmpb_fail(KMPB_EFORMAT, "truncated dump file; tried to read %u file header bytes, only got %zu",
                         PCAP_GLOBAL_HDR, size);
    }
    const uint8_t *file = mmap(NULL, size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (file == MAP_FAILED) return kmpb_fail(KMPB_EIO, "%s: mmap: %s", path, strerror(errno));
    madvise((void *)file, size, MADV_SEQUENTIAL);

    uint32_t magic;
    memcpy(&magic, file, 4);
    int swapped;
    if (magic == 0xa1b2c3d4u || magic == 0xa1b23c4du) swapped = 0;        /* usec / nsec, host order */
    else if (magic == 0xd4c3b2a1u || magic == 0x4d3cb2a1u) swapped = 1;   /* written on the other endianness */
    else {
        munmap((void *)file, size);
        return kmpb_fail(KMPB_EFORMAT, "unknown file format");
    }
    extract_fn extract = proto == KMPB_PROTO_TCP ? kmpb_extract_tcp : kmpb_extract_udp;

    uint64_t n = 0, frames = 0, total = 0;
    walk(file, size, swapped, extract, NULL, NULL, &n, &frames, &total);

    size_t bytes_sz = (size_t)total + CSR_TAIL_PAD, off_sz = (size_t)(n + 1) * sizeof(uint64_t);
    uint8_t *bytes = pinned ? kmpb_host_alloc(bytes_sz) : malloc(bytes_sz);
    uint64_t *offsets = pinned ? kmpb_host_alloc(off_sz) : malloc(off_sz);
    if (bytes == NULL || offsets == NULL) {
        if (pinned) { kmpb_host_free(bytes); kmpb_host_free(offsets); }
        else { free(bytes); free(offsets); }
        munmap((void *)file, size);
        return kmpb_fail(KMPB_ENOMEM, "cannot allocate %zu bytes of %s memory for the payload batch",
                         bytes_sz + off_sz, pinned ? "pinned" : "host");
    }
    walk(file, size, swapped, extract, bytes, offsets, &n, &frames, &total);
    memset(bytes + total, 0, CSR_TAIL_PAD);
    munmap((void *)file, size);

    out->bytes = bytes;
    out->offsets = offsets;
    out->n_packets = n;
    out->n_frames = frames;
    out->total_bytes = total;
    out->pinned = pinned ? 1 : 0;
    return KMPB_OK;
}

void kmpb_free_csr(kmpb_csr *csr)
{
    if (csr == NULL) return;
    if (csr->pinned) { kmpb_host_free(csr->bytes); kmpb_host_free(csr->offsets); }
    else { free(csr->bytes); free(csr->offsets); }
    memset(csr, 0, sizeof *csr);
}

/* pcap_csr.c -- savefile ingest: the step before the hot path (serial.c:91-141, openmp_data.c:94-147).
 *
 * The reference lets libpcap frame the records (pcap_open_offline serial.c:91, pcap_next_ex :115),
 * copies every frame, extracts its payload and stores one malloc'd buffer per payload.  Here the
 * savefile is mapped, the records are framed and their payloads located in one sequential pass, and
 * the accepted payloads are then packed back to back by all host threads (OpenMP, like the
 * reference's own extraction loop at openmp_data.c:128-147) into ONE flat buffer plus an offsets
 * array -- the CSR batch the device consumes -- in pinned memory so the H2D copies can run
 * asynchronously.
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

/* Where an accepted payload lies in the mapped file. */
typedef struct {
    uint64_t at;  /* file offset of the payload's first byte */
    uint32_t len;
} payload_ref;

/* Pass 1, sequential (record n+1 starts where record n's header says): frame the records, run the
 * extractor on every frame and remember where the accepted payloads are. */
static int index_records(const uint8_t *file, size_t size, int swapped, extract_fn extract, payload_ref **refs_out,
                         uint64_t *n_packets, uint64_t *n_frames, uint64_t *total)
{
    size_t at = PCAP_GLOBAL_HDR, cap = 1u << 16;
    uint64_t packets = 0, frames = 0, bytes = 0;
    payload_ref *refs = malloc(cap * sizeof *refs);
    if (!refs) return KMPB_ENOMEM;
    while (size - at >= PCAP_RECORD_HDR) {
        uint32_t caplen = load32(file + at + 8, swapped);
        at += PCAP_RECORD_HDR;
        if (caplen > size - at) break; /* truncated record: libpcap reports an error, the loop ends */
        uint32_t off, len;
        frames++;
        if (extract(file + at, caplen, &off, &len)) {
            if (packets == cap) {
                payload_ref *grown = realloc(refs, 2 * cap * sizeof *refs);
                if (!grown) { free(refs); return KMPB_ENOMEM; }
                refs = grown;
                cap *= 2;
            }
            refs[packets].at = at + off;
            refs[packets].len = len;
            bytes += len;
            packets++;
        }
        at += caplen;
    }
    *refs_out = refs;
    *n_packets = packets;
    *n_frames = frames;
    *total = bytes;
    return KMPB_OK;
}

/* Pass 2, parallel: offsets by a two-level prefix sum, payloads copied by all host threads (the copy
 * is what the reference does per packet at serial.c:124-137, here into one flat buffer). */
static void pack_payloads(const uint8_t *file, const payload_ref *refs, uint64_t n, uint8_t *dst, uint64_t *offsets)
{
    enum { BLOCK = 4096 };
    const uint64_t n_blocks = (n + BLOCK - 1) / BLOCK;
    uint64_t *block_base = malloc((size_t)(n_blocks + 1) * sizeof *block_base);
    if (block_base == NULL) { /* no scratch: plain sequential pack */
        uint64_t bytes = 0;
        for (uint64_t k = 0; k < n; k++) {
            offsets[k] = bytes;
            memcpy(dst + bytes, file + refs[k].at, refs[k].len);
            bytes += refs[k].len;
        }
        offsets[n] = bytes;
        return;
    }
#pragma omp parallel for schedule(static)
    for (uint64_t b = 0; b < n_blocks; b++) {
        uint64_t sum = 0, hi = (b + 1) * BLOCK < n ? (b + 1) * BLOCK : n;
        for (uint64_t k = b * BLOCK; k < hi; k++) sum += refs[k].len;
        block_base[b + 1] = sum;
    }
    block_base[0] = 0;
    for (uint64_t b = 0; b < n_blocks; b++) block_base[b + 1] += block_base[b];
#pragma omp parallel for schedule(dynamic, 4)
    for (uint64_t b = 0; b < n_blocks; b++) {
        uint64_t bytes = block_base[b], hi = (b + 1) * BLOCK < n ? (b + 1) * BLOCK : n;
        for (uint64_t k = b * BLOCK; k < hi; k++) {
            offsets[k] = bytes;
            memcpy(dst + bytes, file + refs[k].at, refs[k].len);
            bytes += refs[k].len;
        }
    }
    offsets[n] = block_base[n_blocks];
    free(block_base);
}

struct kmpb_pcap {
    const uint8_t *file;
    size_t size;
    payload_ref *refs;
    uint64_t n_packets, n_frames, total_bytes;
};

int kmpb_pcap_open(const char *path, int proto, kmpb_pcap **out)
{
    if (path == NULL || out == NULL) return kmpb_fail(KMPB_EINVAL, "kmpb_pcap_open: NULL argument");
    *out = NULL;
    if (proto != KMPB_PROTO_UDP && proto != KMPB_PROTO_TCP) return kmpb_fail(KMPB_EINVAL, "unknown protocol %d", proto);
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
    madvise((void *)file, size, MADV_WILLNEED);

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
    kmpb_pcap *pc = calloc(1, sizeof *pc);
    if (pc == NULL || index_records(file, size, swapped, extract, &pc->refs, &pc->n_packets, &pc->n_frames, &pc->total_bytes) != KMPB_OK) {
        free(pc);
        munmap((void *)file, size);
        return kmpb_fail(KMPB_ENOMEM, "out of memory indexing %s", path);
    }
    pc->file = file;
    pc->size = size;
    *out = pc;
    return KMPB_OK;
}

void kmpb_pcap_close(kmpb_pcap *pc)
{
    if (pc == NULL) return;
    free(pc->refs);
    if (pc->file) munmap((void *)pc->file, pc->size);
    free(pc);
}

uint64_t kmpb_pcap_packets(const kmpb_pcap *pc) { return pc ? pc->n_packets : 0; }
uint64_t kmpb_pcap_frames(const kmpb_pcap *pc) { return pc ? pc->n_frames : 0; }
uint64_t kmpb_pcap_bytes(const kmpb_pcap *pc) { return pc ? pc->total_bytes : 0; }

uint64_t kmpb_pcap_chunk_end(const kmpb_pcap *pc, uint64_t first, uint64_t last, uint64_t max_bytes, uint64_t max_packets,
                             uint64_t *bytes_out)
{
    uint64_t k = first, bytes = 0;
    while (k < last && k - first < max_packets && (k == first || bytes + pc->refs[k].len <= max_bytes)) bytes += pc->refs[k++].len;
    if (bytes_out) *bytes_out = bytes;
    return k;
}

void kmpb_pcap_pack(const kmpb_pcap *pc, uint64_t first, uint64_t count, uint8_t *dst, uint64_t *offsets)
{
    pack_payloads(pc->file, pc->refs + first, count, dst, offsets);
}

int kmpb_load_pcap_csr(const char *path, int proto, int pinned, kmpb_csr *out)
{
    if (path == NULL || out == NULL) return kmpb_fail(KMPB_EINVAL, "kmpb_load_pcap_csr: NULL argument");
    memset(out, 0, sizeof *out);
    kmpb_pcap *pc = NULL;
    int rc = kmpb_pcap_open(path, proto, &pc);
    if (rc != KMPB_OK) return rc;
    const uint64_t n = pc->n_packets, total = pc->total_bytes;
    size_t bytes_sz = (size_t)total + CSR_TAIL_PAD, off_sz = (size_t)(n + 1) * sizeof(uint64_t);
    uint8_t *bytes = pinned ? kmpb_host_alloc(bytes_sz) : malloc(bytes_sz);
    uint64_t *offsets = pinned ? kmpb_host_alloc(off_sz) : malloc(off_sz);
    if (bytes == NULL || offsets == NULL) {
        if (pinned) { kmpb_host_free(bytes); kmpb_host_free(offsets); }
        else { free(bytes); free(offsets); }
        kmpb_pcap_close(pc);
        return kmpb_fail(KMPB_ENOMEM, "cannot allocate %zu bytes of %s memory for the payload batch",
                         bytes_sz + off_sz, pinned ? "pinned" : "host");
    }
    pack_payloads(pc->file, pc->refs, n, bytes, offsets);
    memset(bytes + total, 0, CSR_TAIL_PAD);
    out->bytes = bytes;
    out->offsets = offsets;
    out->n_packets = n;
    out->n_frames = pc->n_frames;
    out->total_bytes = total;
    out->pinned = pinned ? 1 : 0;
    kmpb_pcap_close(pc);
    return KMPB_OK;
}

void kmpb_free_csr(kmpb_csr *csr)
{
    if (csr == NULL) return;
    if (csr->pinned) { kmpb_host_free(csr->bytes); kmpb_host_free(csr->offsets); }
    else { free(csr->bytes); free(csr->offsets); }
    memset(csr, 0, sizeof *csr);
}

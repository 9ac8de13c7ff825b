/*
 * pcap_shim.c -- classic-pcap savefile reader behind the libpcap names the reference calls.
 * TEST INFRASTRUCTURE ONLY (see pcap.h in this directory for why it exists).
 */
#include "pcap.h"

#include <errno.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

struct pcap {
    FILE *fp;
    int swapped;       /* file byte order differs from host */
    int nsec;          /* nanosecond-resolution magic */
    uint32_t snaplen;
    uint32_t linktype;
    struct pcap_pkthdr hdr;
    u_char *buf;
    size_t buf_cap;
    char err[PCAP_ERRBUF_SIZE];
};

static uint32_t bswap32(uint32_t v)
{
    return (v >> 24) | ((v >> 8) & 0x0000ff00u) | ((v << 8) & 0x00ff0000u) | (v << 24);
}

pcap_t *pcap_open_offline(const char *fname, char *errbuf)
{
    FILE *fp = fopen(fname, "rb");
    if (fp == NULL) {
        if (errbuf) snprintf(errbuf, PCAP_ERRBUF_SIZE, "%s: %s", fname, strerror(errno));
        return NULL;
    }
    uint32_t gh[6];
    if (fread(gh, 1, sizeof gh, fp) != sizeof gh) {
        if (errbuf) snprintf(errbuf, PCAP_ERRBUF_SIZE, "truncated dump file; tried to read %zu file header bytes", sizeof gh);
        fclose(fp);
        return NULL;
    }
    int swapped = 0, nsec = 0;
    switch (gh[0]) {
    case 0xa1b2c3d4u: break;
    case 0xa1b23c4du: nsec = 1; break;
    case 0xd4c3b2a1u: swapped = 1; break;
    case 0x4d3cb2a1u: swapped = 1; nsec = 1; break;
    default:
        if (errbuf) snprintf(errbuf, PCAP_ERRBUF_SIZE, "unknown file format");
        fclose(fp);
        return NULL;
    }
    pcap_t *p = calloc(1, sizeof *p);
    if (p == NULL) {
        if (errbuf) snprintf(errbuf, PCAP_ERRBUF_SIZE, "out of memory");
        fclose(fp);
        return NULL;
    }
    p->fp = fp;
    p->swapped = swapped;
    p->nsec = nsec;
    p->snaplen = swapped ? bswap32(gh[4]) : gh[4];
    p->linktype = swapped ? bswap32(gh[5]) : gh[5];
    return p;
}

int pcap_next_ex(pcap_t *p, struct pcap_pkthdr **pkt_header, const u_char **pkt_data)
{
    uint32_t rh[4];
    size_t got = fread(rh, 1, sizeof rh, p->fp);
    if (got == 0) return PCAP_ERROR_BREAK;
    if (got != sizeof rh) {
        snprintf(p->err, sizeof p->err, "truncated dump file; tried to read %zu header bytes, only got %zu", sizeof rh, got);
        return PCAP_ERROR;
    }
    if (p->swapped)
        for (int i = 0; i < 4; i++) rh[i] = bswap32(rh[i]);
    p->hdr.ts.tv_sec = rh[0];
    p->hdr.ts.tv_usec = p->nsec ? rh[1] / 1000 : rh[1];
    p->hdr.caplen = rh[2];
    p->hdr.len = rh[3];
    if (rh[2] > (64u << 20)) {
        snprintf(p->err, sizeof p->err, "invalid packet capture length %u", rh[2]);
        return PCAP_ERROR;
    }
    /* The reference copies header->len bytes out of this buffer (serial.c:117-118); keep it at
     * least that large (zero filled) so caplen < len does not read outside the allocation. */
    size_t need = rh[2] > rh[3] ? rh[2] : rh[3];
    if (need > (64u << 20)) need = rh[2];
    if (need + 16 > p->buf_cap) {
        size_t cap = need + 16 + (need >> 1);
        u_char *nb = realloc(p->buf, cap);
        if (nb == NULL) {
            snprintf(p->err, sizeof p->err, "out of memory");
            return PCAP_ERROR;
        }
        p->buf = nb;
        p->buf_cap = cap;
    }
    memset(p->buf, 0, need + 16);
    if (fread(p->buf, 1, rh[2], p->fp) != rh[2]) {
        snprintf(p->err, sizeof p->err, "truncated dump file; tried to read %u captured bytes", rh[2]);
        return PCAP_ERROR;
    }
    *pkt_header = &p->hdr;
    *pkt_data = p->buf;
    return 1;
}

void pcap_close(pcap_t *p)
{
    if (p == NULL) return;
    if (p->fp) fclose(p->fp);
    free(p->buf);
    free(p);
}

char *pcap_geterr(pcap_t *p) { return p->err; }
int pcap_datalink(pcap_t *p) { return (int)p->linktype; }
int pcap_snapshot(pcap_t *p) { return (int)p->snaplen; }

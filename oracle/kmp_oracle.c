/*
 * kmp_oracle.c -- CPU restatement of the reference's KMP packet-matching path (plain C + OpenMP).
 *
 * TEST INFRASTRUCTURE ONLY -- see kmp_oracle.h for who may use it and how its parity is pinned.
 * Each function cites the reference lines it restates; none of it is copied from the reference.
 */
#define _GNU_SOURCE
#include "kmp_oracle.h"

#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* KMP (serial.c:190-238)                                                                      */
/* ------------------------------------------------------------------------------------------ */

/* serial.c:217-238.  pi[i] = length of the longest proper border of pattern[0..i]. */
int orc_kmp_prefix(const unsigned char *pattern, int m, int *pi)
{
    if (m <= 0) return -1;
    pi[0] = 0;
    int border = 0;
    for (int i = 1; i < m; i++) {
        while (border > 0 && pattern[i] != pattern[border]) border = pi[border - 1];
        if (pattern[i] == pattern[border]) border++;
        pi[i] = border;
    }
    return 0;
}

/* serial.c:190-215.  The reference advances i/j in a hand-rolled loop; this is the same automaton
 * written as "fall back along pi until the next byte extends the match".  After a full match the
 * state drops to pi[m-1] (serial.c:203-206), which is what makes overlapping hits count. */
int64_t orc_kmp_count(const unsigned char *text, int64_t n, const unsigned char *pattern, int m, const int *pi)
{
    if (m <= 0 || n < m) return 0; /* serial.c:193-194 */
    int64_t hits = 0;
    int j = 0;
    for (int64_t i = 0; i < n; i++) {
        unsigned char c = text[i];
        while (j > 0 && pattern[j] != c) j = pi[j - 1];
        if (pattern[j] == c) j++;
        if (j == m) {
            hits++;
            j = pi[m - 1];
        }
    }
    return hits;
}

/* serial.c:191: strlen(text) on a buffer of payload_len bytes with no terminator. */
int64_t orc_text_len(const unsigned char *payload, int64_t payload_len)
{
    if (payload_len <= 0) return 0;
    const unsigned char *z = memchr(payload, 0, (size_t)payload_len);
    return z ? (int64_t)(z - payload) : payload_len;
}

/* ------------------------------------------------------------------------------------------ */
/* Payload extraction (packet_dumping.h:87-188)                                                */
/* ------------------------------------------------------------------------------------------ */

enum { ETH_HDR = 14, IP_MIN_HDR = 20, UDP_HDR = 8, TCP_MIN_HDR = 20, IPPROTO_UDP_ = 17 };

/* packet_dumping.h:87-139.  Five gates, in the reference's order; no ethertype / IP version / IHL>=5
 * / UDP length checks.  The final "+= sizeof(udp_h)" at :133 is sizeof(pointer) == 8 on LP64. */
int orc_udp_payload(const unsigned char *frame, uint32_t frame_len, uint32_t *off, uint32_t *plen)
{
    uint32_t rest = frame_len;
    if (rest < ETH_HDR) return 0;                    /* :94-97   */
    rest -= ETH_HDR;
    if (rest < IP_MIN_HDR) return 0;                 /* :102-105 */
    uint32_t ihl = (uint32_t)(frame[ETH_HDR] & 0x0f) * 4u; /* :107-108 */
    if (rest < ihl) return 0;                        /* :110-113 */
    if (frame[ETH_HDR + 9] != IPPROTO_UDP_) return 0; /* :116-119 */
    rest -= ihl;
    if (rest < UDP_HDR) return 0;                    /* :125-128 */
    *off = ETH_HDR + ihl + UDP_HDR;
    *plen = rest - UDP_HDR;                          /* :133-136 */
    return 1;
}

/* packet_dumping.h:150-188.  No protocol check at all; IHL < 5 or data offset < 5 -> NULL.  Frames
 * shorter than the headers they announce make the reference read outside the frame and wrap an
 * unsigned length (:162,172,182); those frames are skipped here (defined behaviour). */
int orc_tcp_payload(const unsigned char *frame, uint32_t frame_len, uint32_t *off, uint32_t *plen)
{
    if (frame_len < ETH_HDR + 1) return 0;
    uint32_t size_ip = (uint32_t)(frame[ETH_HDR] & 0x0f) * 4u;   /* :165 */
    if (size_ip < IP_MIN_HDR) return 0;                          /* :166-169 */
    uint32_t tcp_at = ETH_HDR + size_ip;
    if (frame_len < tcp_at + 13) return 0;
    uint32_t size_tcp = (uint32_t)((frame[tcp_at + 12] & 0xf0) >> 4) * 4u; /* :175 */
    if (size_tcp < TCP_MIN_HDR) return 0;                        /* :176-179 */
    if (frame_len < tcp_at + size_tcp) return 0;
    *off = tcp_at + size_tcp;
    *plen = frame_len - *off;                                    /* :181-184 */
    return 1;
}

/* ------------------------------------------------------------------------------------------ */
/* Pattern loader (serial.c:54-87)                                                             */
/* ------------------------------------------------------------------------------------------ */

static int is_scanf_space(int c)
{
    return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r';
}

int orc_load_patterns(const char *path, unsigned char **blob_out, uint32_t **off_out, uint32_t *n_out)
{
    FILE *fp = fopen(path, "rb");
    if (fp == NULL) return -1;
    size_t blob_cap = 1024, off_cap = 64, blob_len = 0, n = 0;
    unsigned char *blob = malloc(blob_cap);
    uint32_t *off = malloc(off_cap * sizeof *off);
    int c, rc = 0;
    size_t tok = 0;
    off[0] = 0;
    for (;;) {
        c = fgetc(fp);
        if (c == EOF || is_scanf_space(c)) {
            if (tok > 0) { /* token finished: fscanf("%s") returns one word, serial.c:66 */
                if (n + 2 > off_cap) off = realloc(off, (off_cap *= 2) * sizeof *off);
                off[++n] = (uint32_t)blob_len;
                tok = 0;
            }
            if (c == EOF) break;
            continue;
        }
        if (c == 0) { rc = -3; break; }
        if (++tok > ORC_MAX_PATTERN_LEN) { rc = -2; break; } /* would overflow char str[100], :64 */
        if (blob_len + 1 > blob_cap) blob = realloc(blob, blob_cap *= 2);
        blob[blob_len++] = (unsigned char)c;
    }
    fclose(fp);
    if (rc != 0) {
        free(blob);
        free(off);
        return rc;
    }
    *blob_out = blob;
    *off_out = off;
    *n_out = (uint32_t)n;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* pcap ingest (serial.c:91-141; record framing is libpcap's job in the reference)             */
/* ------------------------------------------------------------------------------------------ */

static uint32_t rd32(const unsigned char *p, int swapped)
{
    uint32_t v;
    memcpy(&v, p, 4);
    if (swapped) v = (v >> 24) | ((v >> 8) & 0xff00u) | ((v << 8) & 0xff0000u) | (v << 24);
    return v;
}

int orc_load_pcap_csr(const char *path, int proto, unsigned char **bytes_out, uint64_t **off_out,
                      uint64_t *n_packets, uint64_t *n_frames)
{
    FILE *fp = fopen(path, "rb");
    if (fp == NULL) return -1;
    unsigned char gh[24];
    if (fread(gh, 1, 24, fp) != 24) { fclose(fp); return -2; }
    uint32_t magic;
    memcpy(&magic, gh, 4);
    int swapped;
    if (magic == 0xa1b2c3d4u || magic == 0xa1b23c4du) swapped = 0;
    else if (magic == 0xd4c3b2a1u || magic == 0x4d3cb2a1u) swapped = 1;
    else { fclose(fp); return -2; }

    size_t bytes_cap = 1 << 16, off_cap = 1 << 10, frame_cap = 1 << 16;
    unsigned char *bytes = malloc(bytes_cap), *frame = malloc(frame_cap);
    uint64_t *off = malloc(off_cap * sizeof *off);
    uint64_t n = 0, frames = 0, total = 0;
    int rc = 0;
    off[0] = 0;
    for (;;) {
        unsigned char rh[16];
        size_t got = fread(rh, 1, 16, fp);
        if (got == 0) break;
        /* a damaged tail makes pcap_next_ex return -1, which ends the reference's
         * `while (... >= 0)` loop (serial.c:115) with the packets read so far */
        if (got != 16) break;
        uint32_t caplen = rd32(rh + 8, swapped);
        if (caplen > (64u << 20)) break;
        if (caplen > frame_cap) frame = realloc(frame, frame_cap = caplen + (caplen >> 1));
        if (fread(frame, 1, caplen, fp) != caplen) break;
        frames++;
        uint32_t poff = 0, plen = 0;
        int ok = proto == ORC_PROTO_TCP ? orc_tcp_payload(frame, caplen, &poff, &plen)
                                        : orc_udp_payload(frame, caplen, &poff, &plen);
        if (!ok) continue; /* serial.c:139-141: NULL payload -> packet dropped */
        if (total + plen > bytes_cap) {
            while (total + plen > bytes_cap) bytes_cap *= 2;
            bytes = realloc(bytes, bytes_cap);
        }
        memcpy(bytes + total, frame + poff, plen);
        total += plen;
        if (n + 2 > off_cap) off = realloc(off, (off_cap *= 2) * sizeof *off);
        off[++n] = total;
    }
    fclose(fp);
    free(frame);
    if (rc != 0) {
        free(bytes);
        free(off);
        return rc;
    }
    *bytes_out = bytes;
    *off_out = off;
    *n_packets = n;
    if (n_frames) *n_frames = frames;
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* Driver (serial.c:148-155, openmp_data.c:149-175) and report (serial.c:163-168)              */
/* ------------------------------------------------------------------------------------------ */

void orc_count_csr(const unsigned char *bytes, const uint64_t *offsets, uint64_t n_packets,
                   const unsigned char *pat_blob, const uint32_t *pat_off, uint32_t n_pat,
                   int64_t *counts, int threads)
{
    memset(counts, 0, (size_t)n_pat * sizeof *counts);
    if (n_pat == 0) return;
    /* one failure table per pattern, built once (serial.c:150-152) */
    int **pi = malloc((size_t)n_pat * sizeof *pi);
    for (uint32_t p = 0; p < n_pat; p++) {
        int m = (int)(pat_off[p + 1] - pat_off[p]);
        pi[p] = malloc((size_t)(m > 0 ? m : 1) * sizeof(int));
        orc_kmp_prefix(pat_blob + pat_off[p], m, pi[p]);
    }
    if (threads < 1) threads = 1;
#pragma omp parallel num_threads(threads)
    {
        /* thread-private counts merged at the end, the shape of openmp_data.c:157-175 */
        int64_t *mine = calloc(n_pat, sizeof *mine);
#pragma omp for schedule(dynamic, 64)
        for (int64_t k = 0; k < (int64_t)n_packets; k++) {
            const unsigned char *payload = bytes + offsets[k];
            int64_t n = orc_text_len(payload, (int64_t)(offsets[k + 1] - offsets[k]));
            for (uint32_t p = 0; p < n_pat; p++)
                mine[p] += orc_kmp_count(payload, n, pat_blob + pat_off[p],
                                         (int)(pat_off[p + 1] - pat_off[p]), pi[p]);
        }
#pragma omp critical
        for (uint32_t p = 0; p < n_pat; p++) counts[p] += mine[p];
        free(mine);
    }
    for (uint32_t p = 0; p < n_pat; p++) free(pi[p]);
    free(pi);
}

char *orc_format_report(const unsigned char *pat_blob, const uint32_t *pat_off, uint32_t n_pat,
                        const int64_t *counts)
{
    static const char head[] =
        "Printing the number of appereances of each string throughout the entire pcap file:\n";
    size_t cap = sizeof head + (size_t)n_pat * (ORC_MAX_PATTERN_LEN + 40), len = 0;
    char *out = malloc(cap);
    len += (size_t)sprintf(out + len, "%s", head);
    for (uint32_t p = 0; p < n_pat; p++) {
        if (counts[p] == 0) continue; /* serial.c:165 */
        uint32_t m = pat_off[p + 1] - pat_off[p];
        memcpy(out + len, pat_blob + pat_off[p], m);
        len += m;
        len += (size_t)sprintf(out + len, ": %d times!\n", (int)counts[p]); /* %d of a C int, :166 */
    }
    out[len] = 0;
    return out;
}

void orc_free(void *p) { free(p); }

int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------------------------ */
/* CLI with the reference's surface (serial.c:24-51), used to diff against oracle/_ref/serial   */
/* ------------------------------------------------------------------------------------------ */
#ifdef ORACLE_MAIN
#include <sys/time.h>
int main(int argc, char **argv)
{
    int proto = ORC_PROTO_UDP, threads = 1;
    const char *usage = "USAGE: ./kmp_oracle <file.pcap> <string.txt> [threads] [tcp/udp]\n";
    if (argc < 3 || argc > 5) { fputs(usage, stdout); return 1; }
    for (int a = 3; a < argc; a++) {
        if (strcmp(argv[a], "udp") == 0) proto = ORC_PROTO_UDP;
        else if (strcmp(argv[a], "tcp") == 0) proto = ORC_PROTO_TCP;
        else if (a == 3 && atoi(argv[a]) > 0) threads = atoi(argv[a]);
        else { fputs(usage, stdout); return 1; }
    }
    unsigned char *blob, *bytes;
    uint32_t *pat_off, n_pat;
    uint64_t *offsets, n_packets, n_frames;
    if (orc_load_patterns(argv[2], &blob, &pat_off, &n_pat) != 0) {
        perror("error opening file: ");
        return 1;
    }
    struct timeval t0, t1;
    gettimeofday(&t0, NULL);
    int rc = orc_load_pcap_csr(argv[1], proto, &bytes, &offsets, &n_packets, &n_frames);
    if (rc != 0) {
        fprintf(stderr, "error reading pcap file: %s\n", rc == -1 ? strerror(errno) : "bad savefile");
        return 1;
    }
    int64_t *counts = calloc(n_pat ? n_pat : 1, sizeof *counts);
    orc_count_csr(bytes, offsets, n_packets, blob, pat_off, n_pat, counts, threads);
    gettimeofday(&t1, NULL);
    char *rep = orc_format_report(blob, pat_off, n_pat, counts);
    fputs(rep, stdout);
    printf("Elapsed time = %f seconds\n",
           (double)(t1.tv_sec - t0.tv_sec) + (double)(t1.tv_usec - t0.tv_usec) / 1e6);
    return 0;
}
#endif

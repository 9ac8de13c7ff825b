/*
 * kmp_oracle.h -- CPU restatement of the reference's KMP packet-matching path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it, and
 * only as the checker or the timed CPU baseline.  The product (libkmpb200.so) never links,
 * loads or calls it and has no CPU fallback.
 *
 * Parity is PINNED: this restatement is checked (tests/test_oracle.py) against
 *   - the outputs of the unmodified reference programs compiled into oracle/_ref/ (serial.c,
 *     openmp_data.c) on every bundled pcap -- committed as tests/golden/expected/ (one .txt per run),
 *   - the reference's own kmp_prefix / kmp_matcher (serial.c:190-238) linked out of serial.c and
 *     run on random inputs -- committed as tests/golden/kmp_vectors.json,
 * by the generator scripts tests/golden/make_golden.py and oracle/Makefile.
 *
 * Defined semantics where the reference is undefined behaviour:
 *   - text of a packet = payload[0 .. min(first NUL byte, payload_len))   (serial.c:191 strlen on
 *     an unterminated malloc(payload_len) buffer, serial.c:125-127),
 *   - occurrences are counted with overlaps (serial.c:203-206),
 *   - frames are read with their captured length (openmp_data.c:114-116); serial.c:117-120 uses the
 *     wire length, identical whenever caplen == len (true for every bundled pcap),
 *   - TCP frames too short for the headers they announce are skipped (the reference reads outside
 *     the frame, packet_dumping.h:162-182).
 */
#ifndef KMP_ORACLE_H
#define KMP_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_PROTO_UDP 0
#define ORC_PROTO_TCP 1
#define ORC_MAX_PATTERN_LEN 99 /* char str[100] + fscanf("%s"): serial.c:64-66 */

/* serial.c:217-238 kmp_prefix.  pi must hold m ints.  Returns 0, or -1 if m <= 0. */
int orc_kmp_prefix(const unsigned char *pattern, int m, int *pi);

/* serial.c:190-215 kmp_matcher with explicit lengths: overlapping occurrences of pattern[0..m)
 * in text[0..n).  The caller applies the NUL rule (orc_text_len) first. */
int64_t orc_kmp_count(const unsigned char *text, int64_t n, const unsigned char *pattern, int m, const int *pi);

/* strlen() of an unterminated payload, made defined: min(first NUL, payload_len). serial.c:191. */
int64_t orc_text_len(const unsigned char *payload, int64_t payload_len);

/* packet_dumping.h:87-139 dump_UDP_packet.  Returns 1 and sets *off / *plen (payload offset inside
 * the frame and payload length), or 0 when the reference returns NULL. */
int orc_udp_payload(const unsigned char *frame, uint32_t frame_len, uint32_t *off, uint32_t *plen);

/* packet_dumping.h:150-188 dump_TCP_packet, same convention. */
int orc_tcp_payload(const unsigned char *frame, uint32_t frame_len, uint32_t *off, uint32_t *plen);

/* serial.c:54-87 pattern loader: whitespace-separated tokens of at most 99 bytes, duplicates kept,
 * file order.  Allocates *blob (concatenated tokens) and *pat_off (n+1 offsets); caller frees with
 * orc_free.  Returns 0; -1 cannot open (errno set); -2 token too long; -3 NUL byte in file. */
int orc_load_patterns(const char *path, unsigned char **blob, uint32_t **pat_off, uint32_t *n_pat);

/* serial.c:91-141 ingest: classic pcap -> flat CSR of accepted payloads (bytes + n+1 offsets).
 * Returns 0; -1 cannot open; -2 not a classic pcap.  A truncated trailing record ends the walk like
 * the reference's `>= 0` loop does.  *n_frames = records read. */
int orc_load_pcap_csr(const char *path, int proto, unsigned char **bytes, uint64_t **offsets,
                      uint64_t *n_packets, uint64_t *n_frames);

/* serial.c:148-155 driver (openmp_data.c:157-175 when threads > 1): counts[i] = sum over packets
 * of orc_kmp_count(text(packet), pattern i).  counts has n_pat entries and is overwritten. */
void orc_count_csr(const unsigned char *bytes, const uint64_t *offsets, uint64_t n_packets,
                   const unsigned char *pat_blob, const uint32_t *pat_off, uint32_t n_pat,
                   int64_t *counts, int threads);

/* serial.c:163-168 report body (without the Elapsed line) into a malloc'd string. */
char *orc_format_report(const unsigned char *pat_blob, const uint32_t *pat_off, uint32_t n_pat,
                        const int64_t *counts);

void orc_free(void *p);
int orc_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif

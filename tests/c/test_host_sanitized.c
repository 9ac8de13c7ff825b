/* test_host_sanitized.c -- the product's host side (pattern loader, savefile reader, payload extraction, CSR packer,
 * table builder, report printer) run over every file given on the command line, meant to be built with
 * -fsanitize=address,undefined (tests/test_host.py builds and runs it; the reference's own loader is not
 * sanitizer-clean).  No GPU: kmpb_host_alloc/kmpb_host_free are stubbed with malloc/free.
 *
 *   test_host_sanitized <strings.txt> <file.pcap>...
 * prints "<file> <proto> packets=<n> bytes=<n> nulfree=<n>" per file and protocol. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kmpb200.h"
#include "kmpb_internal.h"

void *kmpb_host_alloc(size_t n) { return malloc(n ? n : 1); }
void kmpb_host_free(void *p) { free(p); }

int main(int argc, char **argv)
{
    if (argc < 3) return 2;
    kmpb_patterns pats;
    if (kmpb_load_patterns_file(argv[1], &pats) != KMPB_OK) { fprintf(stderr, "patterns: %s\n", kmpb_last_error()); return 1; }
    kmpb_tables t;
    if (kmpb_tables_build(&t, pats.blob, pats.pat_off, pats.n_pat) != KMPB_OK) { fprintf(stderr, "tables: %s\n", kmpb_last_error()); return 1; }
    uint64_t *counts = calloc(pats.n_pat ? pats.n_pat : 1, sizeof *counts);
    for (uint32_t i = 0; i < pats.n_pat; i++) counts[i] = i % 3; /* some lines print, some do not */
    FILE *sink = fopen("/dev/null", "w");
    kmpb_print_report(sink, &pats, counts);
    fclose(sink);
    int bad = 0;
    for (int a = 2; a < argc; a++) {
        for (int proto = 0; proto < 2; proto++) {
            kmpb_csr csr;
            int rc = kmpb_load_pcap_csr(argv[a], proto == 0 ? KMPB_PROTO_UDP : KMPB_PROTO_TCP, 0, &csr);
            if (rc != KMPB_OK) { printf("%s %s error %d\n", argv[a], proto ? "tcp" : "udp", rc); continue; }
            uint64_t nulfree = 0;
            for (uint64_t k = 0; k < csr.n_packets; k++) {
                if (csr.offsets[k + 1] < csr.offsets[k] || csr.offsets[k + 1] > csr.total_bytes) { bad = 1; break; }
                nulfree += memchr(csr.bytes + csr.offsets[k], 0, csr.offsets[k + 1] - csr.offsets[k]) == NULL;
            }
            printf("%s %s packets=%llu bytes=%llu nulfree=%llu\n", argv[a], proto ? "tcp" : "udp",
                   (unsigned long long)csr.n_packets, (unsigned long long)csr.total_bytes, (unsigned long long)nulfree);
            kmpb_free_csr(&csr);
            /* the streamed reader over the same file, in small chunks */
            kmpb_pcap *pc = NULL;
            if (kmpb_pcap_open(argv[a], proto == 0 ? KMPB_PROTO_UDP : KMPB_PROTO_TCP, &pc) == KMPB_OK) {
                uint64_t n = kmpb_pcap_packets(pc), first = 0, total = 0;
                while (first < n) {
                    uint64_t bytes = 0;
                    uint64_t end = kmpb_pcap_chunk_end(pc, first, n, 4096, 7, &bytes);
                    uint8_t *dst = malloc(bytes + 1);
                    uint64_t *off = malloc((end - first + 1) * sizeof *off);
                    kmpb_pcap_pack(pc, first, end - first, dst, off);
                    if (off[end - first] != bytes) bad = 1;
                    total += bytes;
                    free(dst); free(off);
                    first = end;
                }
                if (total != kmpb_pcap_bytes(pc)) bad = 1;
                kmpb_pcap_close(pc);
            }
        }
    }
    free(counts);
    kmpb_tables_free(&t);
    kmpb_free_patterns(&pats);
    return bad;
}

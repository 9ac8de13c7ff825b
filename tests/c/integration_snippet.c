/* integration_snippet.c -- the call sequence of INTEGRATION.md section 1 as a program of its own: what a maintainer of
 * serial.c writes in place of serial.c:148-155 (failure tables + packet x pattern loop).  Built and run by
 * tests/test_host.py: it must compile and link against libkmpb200.so with nothing but include/kmpb200.h; without a
 * B200 it must fail at kmpb_create with the library's message (there is no CPU fallback), with one it prints the
 * counts of two patterns in three payloads. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kmpb200.h"

int main(void)
{
    /* what serial.c has at line 146: char **array_of_strings, char **array_of_payloads (+ the lengths it drops) */
    const char *array_of_strings[] = {"aa", "http"};
    const char *array_of_payloads[] = {"aaaa http", "xhttphttp\0http", "a"};
    const unsigned payload_lens[] = {9, 14, 1};
    const int array_of_strings_length = 2, count = 3;

    uint32_t n_pat = (uint32_t)array_of_strings_length, *pat_off = malloc((n_pat + 1) * sizeof *pat_off);
    size_t blob_len = 0;
    for (uint32_t i = 0; i < n_pat; i++) { pat_off[i] = (uint32_t)blob_len; blob_len += strlen(array_of_strings[i]); }
    pat_off[n_pat] = (uint32_t)blob_len;
    uint8_t *blob = malloc(blob_len ? blob_len : 1);
    for (uint32_t i = 0; i < n_pat; i++) memcpy(blob + pat_off[i], array_of_strings[i], pat_off[i + 1] - pat_off[i]);

    kmpb_ctx *ctx;
    if (kmpb_create(&ctx, 0) || kmpb_set_patterns(ctx, blob, pat_off, n_pat)) {
        fprintf(stderr, "kmpb: %s\n", kmpb_last_error());
        return 1;
    }
    uint64_t *offsets = malloc((count + 1) * sizeof *offsets), total = 0;
    for (int k = 0; k < count; k++) { offsets[k] = total; total += payload_lens[k]; }
    offsets[count] = total;
    uint8_t *bytes = malloc(total + 64);
    for (int k = 0; k < count; k++) memcpy(bytes + offsets[k], array_of_payloads[k], payload_lens[k]);
    uint64_t counts[2] = {0, 0};
    if (kmpb_count_host(ctx, bytes, offsets, (uint64_t)count, counts)) {
        fprintf(stderr, "kmpb: %s\n", kmpb_last_error());
        return 1;
    }
    /* "aa" in "aaaa" overlaps: 3 (serial.c:203-206); the text of payload 2 ends at its NUL (serial.c:191) */
    for (uint32_t i = 0; i < n_pat; i++) printf("%s: %d times!\n", array_of_strings[i], (int)counts[i]);
    kmpb_destroy(ctx);
    free(bytes); free(offsets); free(blob); free(pat_off);
    return 0;
}

/* patterns.c -- the pattern set as the reference reads it (serial.c:54-87): fscanf("%s") tokens,
 * i.e. maximal runs of non-whitespace bytes, in file order, duplicates kept, at most 99 bytes each
 * (char str[100], :64).  The reference overflows its buffer on longer tokens and mis-measures tokens
 * containing NUL; both are rejected here with KMPB_EFORMAT. */
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kmpb_internal.h"

static int scanf_ws(unsigned char c) { return c == ' ' || (c >= '\t' && c <= '\r'); }

int kmpb_load_patterns_file(const char *path, kmpb_patterns *out)
{
    if (path == NULL || out == NULL) return kmpb_fail(KMPB_EINVAL, "kmpb_load_patterns_file: NULL argument");
    memset(out, 0, sizeof *out);
    FILE *fp = fopen(path, "rb");
    if (fp == NULL) {
        int e = errno;
        kmpb_fail(KMPB_EIO, "%s: %s", path, strerror(e));
        errno = e;
        return KMPB_EIO;
    }
    /* slurp: strings files are tiny */
    size_t cap = 1 << 12, len = 0;
    uint8_t *text = malloc(cap);
    if (text == NULL) { fclose(fp); return kmpb_fail(KMPB_ENOMEM, "out of memory"); }
    for (;;) {
        if (len == cap) {
            uint8_t *grown = realloc(text, cap *= 2);
            if (grown == NULL) { free(text); fclose(fp); return kmpb_fail(KMPB_ENOMEM, "out of memory"); }
            text = grown;
        }
        size_t got = fread(text + len, 1, cap - len, fp);
        if (got == 0) break;
        len += got;
    }
    fclose(fp);

    /* pass 1: count tokens and validate */
    uint32_t n = 0;
    size_t run = 0, blob_len = 0;
    for (size_t i = 0; i <= len; i++) {
        if (i == len || scanf_ws(text[i])) {
            if (run) n++;
            run = 0;
            continue;
        }
        if (text[i] == 0) { free(text); return kmpb_fail(KMPB_EFORMAT, "%s: NUL byte in pattern file", path); }
        if (++run > KMPB_MAX_PATTERN_LEN) {
            free(text);
            return kmpb_fail(KMPB_EFORMAT, "%s: pattern longer than %d bytes", path, KMPB_MAX_PATTERN_LEN);
        }
        blob_len++;
    }
    out->blob = malloc(blob_len ? blob_len : 1);
    out->pat_off = malloc(((size_t)n + 1) * sizeof *out->pat_off);
    if (out->blob == NULL || out->pat_off == NULL) {
        free(text);
        kmpb_free_patterns(out);
        return kmpb_fail(KMPB_ENOMEM, "out of memory");
    }
    /* pass 2: copy tokens */
    uint32_t k = 0, at = 0;
    out->pat_off[0] = 0;
    run = 0;
    for (size_t i = 0; i <= len; i++) {
        if (i == len || scanf_ws(text[i])) {
            if (run) out->pat_off[++k] = at;
            run = 0;
            continue;
        }
        out->blob[at++] = text[i];
        run++;
    }
    out->n_pat = n;
    free(text);
    return KMPB_OK;
}

void kmpb_free_patterns(kmpb_patterns *p)
{
    if (p == NULL) return;
    free(p->blob);
    free(p->pat_off);
    memset(p, 0, sizeof *p);
}

/* serial.c:163-168 */
int kmpb_print_report(void *stream, const kmpb_patterns *pats, const uint64_t *counts)
{
    FILE *fp = (FILE *)stream;
    if (fp == NULL || pats == NULL || (pats->n_pat && counts == NULL))
        return kmpb_fail(KMPB_EINVAL, "kmpb_print_report: NULL argument");
    fputs("Printing the number of appereances of each string throughout the entire pcap file:\n", fp);
    for (uint32_t i = 0; i < pats->n_pat; i++) {
        if (counts[i] == 0) continue;
        fwrite(pats->blob + pats->pat_off[i], 1, pats->pat_off[i + 1] - pats->pat_off[i], fp);
        fprintf(fp, ": %d times!\n", (int)counts[i]); /* the reference's counters are C ints */
    }
    return KMPB_OK;
}

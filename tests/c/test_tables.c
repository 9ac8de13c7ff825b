/* test_tables.c -- CPU check of the host-built tables (csrc/host/automaton.c), no GPU involved.
 *
 * For seeded random pattern sets and texts it checks, against a naive overlapping counter:
 *   1. walking the union DFA from the root reports exactly the naive per-pattern counts
 *      (the property that makes the union automaton a drop-in for P independent kmp_matcher calls);
 *   2. the probe tables, probed the way the device probes them, report the same counts;
 *   3. the shift-and prefilter (6-bit fields, two bytes per update) never misses: for every true occurrence starting
 *      at s the filter word after byte s+3 (text padded with zero bytes) has one of bits 18..22 set; its NUL bit says
 *      exactly "the byte three back is NUL"; a report lingers one step; the two-byte update equals two one-byte ones.
 * Exit status 0 = all good.  Built and run by tests/test_host.py.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "kmpb_internal.h"

static uint64_t rng_state = 0x1234567;
static uint32_t rnd(void)
{
    rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
    return (uint32_t)(rng_state >> 11);
}

static int run_case(int n_pat, int max_len, const char *alpha, int text_len, int nul_every)
{
    int na = (int)strlen(alpha);
    uint8_t *blob = malloc((size_t)n_pat * 99 + 1);
    uint32_t *off = malloc(((size_t)n_pat + 1) * sizeof *off);
    uint32_t at = 0;
    for (int i = 0; i < n_pat; i++) {
        off[i] = at;
        int len = 1 + (int)(rnd() % (uint32_t)max_len);
        if (i > 0 && rnd() % 7 == 0) { /* duplicate of an earlier pattern */
            int j = (int)(rnd() % (uint32_t)i);
            len = (int)(off[j + 1] - off[j]);
            memcpy(blob + at, blob + off[j], (size_t)len);
        } else {
            for (int k = 0; k < len; k++) blob[at + k] = (uint8_t)alpha[rnd() % (uint32_t)na];
        }
        at += (uint32_t)len;
    }
    off[n_pat] = at;
    uint8_t *text = calloc((size_t)text_len + 8, 1);
    for (int i = 0; i < text_len; i++) text[i] = (nul_every && rnd() % (uint32_t)nul_every == 0) ? 0 : (uint8_t)alpha[rnd() % (uint32_t)na];
    for (int r = 0; r < 4 && n_pat; r++) { /* plant some patterns so hits exist with large alphabets */
        int p = (int)(rnd() % (uint32_t)n_pat), len = (int)(off[p + 1] - off[p]);
        if (len < text_len) memcpy(text + rnd() % (uint32_t)(text_len - len), blob + off[p], (size_t)len);
    }

    kmpb_tables t;
    if (kmpb_tables_build_ex(&t, blob, off, (uint32_t)n_pat, 1) != 0) { fprintf(stderr, "build failed: %s\n", kmpb_last_error()); return 1; }

    /* 1. DFA counts (NULs are ordinary mismatching bytes here: class 0) */
    uint64_t *got = calloc((size_t)t.n_uniq + 1, sizeof *got);
    uint32_t state = 0;
    for (int i = 0; i < text_len; i++) {
        uint32_t e = t.next[(size_t)state * t.n_class + t.byte_class[text[i]]];
        state = e & 0x7fffffffu;
        if ((e >> 31) != (t.out_head[state + 1] > t.out_head[state])) { fprintf(stderr, "report flag wrong\n"); return 1; }
        for (uint32_t o = t.out_head[state]; o < t.out_head[state + 1]; o++) got[t.out_id[o]]++;
    }
    int bad = 0;
    /* 1b. the bare trie, walked from every start position, reports the same counts */
    uint64_t *got_trie = calloc((size_t)t.n_uniq + 1, sizeof *got_trie);
    for (int s = 0; s < text_len; s++) {
        uint32_t node = 0;
        for (int k = s; k < text_len; k++) {
            uint32_t e = t.trie[(size_t)node * t.n_class + t.byte_class[text[k]]];
            if (!e) break;
            node = e & 0x7fffffffu;
            if ((e >> 31) != (t.state_term[node] != 0xffffffffu)) { fprintf(stderr, "trie terminal flag wrong\n"); return 1; }
            if (e >> 31) got_trie[t.state_term[node]]++;
        }
    }
    for (uint32_t u = 0; u < t.n_uniq; u++)
        if (got_trie[u] != got[u]) { fprintf(stderr, "uniq %u: trie %llu dfa %llu\n", u, (unsigned long long)got_trie[u], (unsigned long long)got[u]); bad = 1; }
    free(got_trie);
    /* 1c. the probe tables, probed at every start position the way the device probes them (one slot of table A by the
     *     first two bytes, one of table B by the first three, masked compare of the first 8 bytes, then the remaining
     *     pattern words), report the same counts */
    if (t.n_uniq) {
        const uint32_t *v = t.vtab;
        uint64_t *got_hash = calloc((size_t)t.n_uniq + 1, sizeof *got_hash);
        uint32_t n_rec = 0;
        for (int tab = 0; tab < 2; tab++) {
            const uint32_t so = v[1 + 2 * tab], shift = v[2 + 2 * tab];
            if (!so) continue;
            for (uint32_t s = 0; s < (1u << (32 - shift)); s++) {
                if (v[so + 2 * s] != n_rec) { fprintf(stderr, "table %d slot %u: chains are not contiguous\n", tab, s); bad = 1; }
                n_rec += v[so + 2 * s + 1];
            }
        }
        for (int s = 0; s < text_len; s++) {
            uint32_t x0 = 0, x1 = 0;
            for (int k = 0; k < 4; k++) x0 |= (uint32_t)text[s + k] << (8 * k); /* text is zero padded */
            for (int k = 0; k < 4; k++) x1 |= (uint32_t)text[s + 4 + k] << (8 * k);
            if (v[5] && v[v[5] + (x0 & 0xffu)] != 0xffffffffu) got_hash[v[v[5] + (x0 & 0xffu)]]++;
            for (int tab = 0; tab < 2; tab++) {
                const uint32_t so = v[1 + 2 * tab], shift = v[2 + 2 * tab];
                if (!so) continue;
                const uint32_t *e = v + so + 2 * kmpb_vtab_slot(x0 & (tab ? 0xffffffu : 0xffffu), shift);
                for (uint32_t r = 0; r < e[1]; r++) {
                    const uint32_t *rec = v + v[6] + 8 * (e[0] + r);
                    const uint32_t len = rec[4], u = rec[5];
                    const uint8_t *pb = (const uint8_t *)(v + v[7] + rec[6]);
                    if (u >= t.n_uniq || len != t.uniq_len[u] || (tab ? len < 3 : len != 2) || memcmp(pb, t.uniq_blob + t.uniq_off[u], len)) { fprintf(stderr, "record of slot: wrong pattern\n"); bad = 1; break; }
                    if (((x0 ^ rec[0]) & rec[1]) | ((x1 ^ rec[2]) & rec[3])) continue;
                    if (s + (int)len > text_len) continue;
                    if (len > 8 && memcmp(text + s + 8, pb + 8, len - 8)) continue;
                    if (memcmp(text + s, pb, len)) { fprintf(stderr, "record masks wrong for uniq %u\n", u); bad = 1; }
                    got_hash[u]++;
                }
            }
        }
        for (uint32_t u = 0; u < t.n_uniq; u++)
            if (got_hash[u] != got[u]) { fprintf(stderr, "uniq %u: hash %llu dfa %llu\n", u, (unsigned long long)got_hash[u], (unsigned long long)got[u]); bad = 1; }
        free(got_hash);
    }
    for (int p = 0; p < n_pat; p++) {
        int len = (int)(off[p + 1] - off[p]);
        uint64_t want = 0;
        for (int i = 0; i + len <= text_len; i++) want += memcmp(text + i, blob + off[p], (size_t)len) == 0;
        uint32_t u = t.pat_to_uniq[p];
        if (t.uniq_len[u] != (uint32_t)len || memcmp(t.uniq_blob + t.uniq_off[u], blob + off[p], (size_t)len)) { fprintf(stderr, "uniq map wrong\n"); return 1; }
        if (got[u] != want) { fprintf(stderr, "pattern %d: dfa %llu naive %llu\n", p, (unsigned long long)got[u], (unsigned long long)want); bad = 1; }
    }
    /* 3. the filter in 6-bit fields (t.filter6: 5 buckets + NUL, depth 4 + a lingering field), as the union kernel uses
     *    it: superset of the true occurrences; bit 23 after the update of byte i says "byte i-3 is NUL" (exactly); the
     *    report of the previous byte lingers in bits 24..29; the two-byte update equals two one-byte updates */
    {
        const uint32_t *w6 = t.filter6;
        const uint32_t all30 = 0x3fffffffu;
        for (int p = 0; p < n_pat; p++) {
            int len = (int)(off[p + 1] - off[p]);
            for (int s0 = 0; s0 + len <= text_len; s0++) {
                if (memcmp(text + s0, blob + off[p], (size_t)len)) continue;
                uint32_t S6 = 0;
                for (int k = 0; k < 4; k++) S6 = ((S6 << 6) | 0x3fu) & w6[text[s0 + k]]; /* text is zero padded */
                if (!(S6 & (0x1fu << 18))) { fprintf(stderr, "filter6 missed pattern %d at %d\n", p, s0); bad = 1; }
            }
        }
        uint32_t S1 = 0, S2 = 0;
        for (int i = 0; i + 1 < text_len; i += 2) {
            const uint32_t b0 = text[i], b1 = text[i + 1];
            const uint32_t mid = ((S1 << 6) | 0x3fu) & w6[b0];
            if (i >= 3 && ((mid >> 23) & 1u) != (text[i - 3] == 0)) { fprintf(stderr, "filter6 NUL bit wrong at %d\n", i); bad = 1; break; }
            S1 = ((mid << 6) | 0x3fu) & w6[b1];
            if (i >= 2 && ((S1 >> 23) & 1u) != (text[i - 2] == 0)) { fprintf(stderr, "filter6 NUL bit wrong at %d\n", i + 1); bad = 1; break; }
            if (((S1 >> 24) & 0x3fu) != ((mid >> 18) & 0x3fu)) { fprintf(stderr, "filter6: report does not linger at %d\n", i); bad = 1; break; }
            S2 = ((S2 << 12) | 0xfffu) & ((w6[b0] << 6) | 0x3fu) & w6[b1];
            if ((S1 & all30) != (S2 & all30)) { fprintf(stderr, "filter6: two-byte update differs at %d\n", i); bad = 1; break; }
        }
    }
    free(got); free(text); free(blob); free(off);
    kmpb_tables_free(&t);
    return bad;
}

int main(void)
{
    int bad = 0;
    const char *alphas[] = {"ab", "abc", "abcdefghijklmnopqrstuvwxyz0123456789", "aA-_ .:/"};
    for (int trial = 0; trial < 200 && !bad; trial++) {
        int n_pat = (int[]){0, 1, 2, 5, 8, 30, 97, 300}[trial % 8];
        int max_len = (int[]){1, 2, 3, 4, 5, 12, 40, 99}[(trial / 8) % 8];
        bad |= run_case(n_pat, max_len, alphas[trial % 4], 50 + (int)(rnd() % 3000), trial % 3 == 0 ? 40 : 0);
    }
    /* limits */
    kmpb_tables t;
    uint8_t longpat[101]; uint32_t off2[2] = {0, 100};
    memset(longpat, 'x', sizeof longpat);
    if (kmpb_tables_build(&t, longpat, off2, 1) != KMPB_EINVAL) { fprintf(stderr, "100-byte pattern accepted\n"); bad = 1; }
    off2[1] = 0;
    if (kmpb_tables_build(&t, longpat, off2, 1) != KMPB_EINVAL) { fprintf(stderr, "empty pattern accepted\n"); bad = 1; }
    longpat[1] = 0; off2[1] = 3;
    if (kmpb_tables_build(&t, longpat, off2, 1) != KMPB_EINVAL) { fprintf(stderr, "NUL pattern accepted\n"); bad = 1; }
    printf(bad ? "FAIL\n" : "OK\n");
    return bad;
}

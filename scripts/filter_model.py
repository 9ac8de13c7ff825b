#!/usr/bin/env python3
"""CPU model of the union engine's prefilter: candidate and event rates of different filter geometries on
the BASELINE workloads (no GPU).  A geometry = (depth, pattern buckets); a bucket's stage d passes the byte
values some member has at depth d (every value when a member is shorter).  Usage: scripts/filter_model.py"""
import sys, os, itertools, random
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multithreading_string_matching_b200 import matcher as M

ALPHA = 96.0


def bucket_cost(members, depth):
    c = 1.0
    for d in range(depth):
        if any(len(p) <= d for p in members):
            continue
        c *= min(1.0, len({p[d] for p in members}) / ALPHA)
    return c if members else 0.0


def optimise(pats, depth, nb, seed=1):
    """contiguous DP split of the sorted patterns, then hill climbing (moves + swaps), as automaton.c does"""
    pats = sorted(set(pats), key=lambda p: (min(len(p), depth), p))
    n = len(pats)
    nb = min(nb, n)
    cost = {}
    for i in range(n):
        for j in range(i + 1, n + 1):
            cost[(i, j)] = bucket_cost(pats[i:j], depth)
    best = [[1e300] * (n + 1) for _ in range(nb + 1)]
    frm = [[0] * (n + 1) for _ in range(nb + 1)]
    best[0][0] = 0.0
    for k in range(1, nb + 1):
        for j in range(n + 1):
            for i in range(k - 1, j):
                if best[k - 1][i] < 1e300:
                    c = best[k - 1][i] + cost[(i, j)]
                    if c < best[k][j]:
                        best[k][j] = c
                        frm[k][j] = i
    cuts = [n]
    j = n
    for k in range(nb, 0, -1):
        j = frm[k][j]
        cuts.append(j)
    cuts = cuts[::-1]
    bucket = [0] * n
    for b in range(nb):
        for i in range(cuts[b], cuts[b + 1]):
            bucket[i] = b

    def total():
        return sum(bucket_cost([pats[i] for i in range(n) if bucket[i] == b], depth) for b in range(nb))
    cur = total()
    improved = True
    sweeps = 0
    while improved and sweeps < 30:
        improved = False
        sweeps += 1
        for i in range(n):
            a = bucket[i]
            bestc, to = cur, a
            for b in range(nb):
                if b == a:
                    continue
                bucket[i] = b
                c = total()
                if c < bestc - 1e-15:
                    bestc, to = c, b
            bucket[i] = to
            if to != a:
                cur = bestc
                improved = True
        for i in range(n):
            for j in range(i + 1, n):
                if bucket[i] == bucket[j]:
                    continue
                bucket[i], bucket[j] = bucket[j], bucket[i]
                c = total()
                if c < cur - 1e-15:
                    cur = c
                    improved = True
                else:
                    bucket[i], bucket[j] = bucket[j], bucket[i]
    return pats, bucket, cur


def stage_tables(pats, bucket, depth, nb):
    """A[d][byte] = bitmask of buckets passing byte at depth d"""
    A = np.zeros((depth, 256), dtype=np.uint32)
    for b in range(nb):
        mem = [p for p, bb in zip(pats, bucket) if bb == b]
        if not mem:
            continue
        for d in range(depth):
            if any(len(p) <= d for p in mem):
                A[d, :] |= 1 << b
            else:
                for p in mem:
                    A[d, p[d]] |= 1 << b
    return A


def rates(text, A, depth):
    n = len(text) - depth
    c = A[0][text[:n]]
    for d in range(1, depth):
        c = c & A[d][text[d:d + n]]
    cand = c != 0
    nul = text[:n] == 0
    rep = cand | nul
    g = n // 32 * 32
    groups = rep[:g].reshape(-1, 32).any(axis=1)
    quarters = rep[:g].reshape(-1, 8).any(axis=1)
    kb = g / 1024.0
    return cand[:g].sum() / kb, nul[:g].sum() / kb, groups.sum() / kb, quarters.sum() / kb


def main():
    pats = M.load_patterns(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "data", "strings.txt"))
    pats = [bytes(p) for p in pats]
    syn = M.Synth(seed=0xB200, payload_len=1400, len_mode=0, plants=2, plant_patterns=pats)
    data, offsets = syn.fill_host(0, 20000)
    text = np.frombuffer(data, dtype=np.uint8)[: int(offsets[-1])]
    rnd = np.random.default_rng(1).integers(0x20, 0x7f, size=len(text), dtype=np.uint8)
    print("C3 text: %d bytes; columns: candidates/KB, NULs/KB, 32-byte groups with a report/KB, 8-byte quarters/KB" % len(text))
    for depth, nb in [(4, 7), (4, 5), (3, 7), (3, 8), (4, 8), (3, 5), (2, 7), (4, 15), (3, 15)]:
        sp, bucket, est = optimise(pats, depth, nb)
        A = stage_tables(sp, bucket, depth, nb)
        r = rates(text, A, depth)
        r2 = rates(rnd, A, depth)
        print("depth %d buckets %2d: estimate %.2e/byte | C3 cand %.2f nul %.2f groups %.2f quarters %.2f | plain printable cand %.2f groups %.2f"
              % (depth, nb, est, r[0], r[1], r[2], r[3], r2[0], r2[2]))


if __name__ == "__main__" and len(sys.argv) == 1:
    main()


def main2():
    """geometries with the two-byte patterns in a chain of their own (exact pairs)"""
    pats = M.load_patterns(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "data", "strings.txt"))
    pats = [bytes(p) for p in pats]
    syn = M.Synth(seed=0xB200, payload_len=1400, len_mode=0, plants=2, plant_patterns=pats)
    data, offsets = syn.fill_host(0, 20000)
    text = np.frombuffer(data, dtype=np.uint8)[: int(offsets[-1])]
    short = sorted({p for p in pats if len(p) <= 2})
    longer = [p for p in pats if len(p) > 2]
    for depth, nb in [(4, 5), (4, 6), (3, 7), (4, 7)]:
        sp, bucket, est = optimise(longer, depth, nb)
        A = stage_tables(sp, bucket, depth, nb)
        # the short chain: bucket bit nb, exact in two bytes, open beyond
        for d in range(depth):
            if d >= 2:
                A[d, :] |= 1 << nb
            else:
                for p in short:
                    if len(p) > d:
                        A[d, p[d]] |= 1 << nb
                    else:
                        A[d, :] |= 1 << nb
        r = rates(text, A, depth)
        print("depth %d buckets %d (+ pair chain): C3 cand %.2f groups %.2f quarters %.2f" % (depth, nb, r[0], r[2], r[3]))


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "2":
    main2()

"""The oracle (oracle/kmp_oracle.c) against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import json
import os
import random

import numpy as np
import pytest

from conftest import DATA, GOLDEN, golden_runs


def naive_count(text, pat):
    """Overlapping occurrences with the NUL rule, independent of KMP."""
    z = text.find(b"\0")
    if z >= 0:
        text = text[:z]
    return sum(1 for i in range(len(text) - len(pat) + 1) if text[i:i + len(pat)] == pat) if pat else 0


@pytest.mark.parametrize("pcap,proto,expected", golden_runs(), ids=lambda v: v if isinstance(v, str) else "")
def test_bundled_pcaps_match_reference_stdout(oracle, strings_txt, pcap, proto, expected):
    data, offsets, _ = oracle.load_pcap_csr(os.path.join(DATA, pcap + ".pcap"), proto)
    patterns = oracle.load_patterns(strings_txt)
    counts = oracle.count_csr(data, offsets, patterns, threads=4)
    assert oracle.format_report(patterns, counts) == expected


def test_known_answers_from_survey(oracle, strings_txt):
    """SURVEY.md section 4 table (serial.c on the bundled pcaps)."""
    patterns = oracle.load_patterns(strings_txt)
    assert len(patterns) == 97 and len(set(patterns)) == 87 and patterns.count(b"ack") == 3
    data, offsets, frames = oracle.load_pcap_csr(os.path.join(DATA, "big_udp.pcap"))
    assert (frames, len(offsets) - 1, int(offsets[-1])) == (3580, 3358, 599424)
    counts = dict(zip(patterns, oracle.count_csr(data, offsets, patterns)))
    assert (counts[b"http"], counts[b"Linux"], counts[b"NOTIFY"], counts[b"ack"], counts[b"port"]) == (879, 407, 704, 8, 12)
    data, offsets, frames = oracle.load_pcap_csr(os.path.join(DATA, "udp_1000.pcap"))
    assert (frames, len(offsets) - 1, int(offsets[-1])) == (1000, 321, 84519)
    data, offsets, frames = oracle.load_pcap_csr(os.path.join(DATA, "very_big_udp.pcap"))
    assert (frames, len(offsets) - 1, int(offsets[-1])) == (13768, 13768, 1321746)
    assert sum(oracle.count_csr(data, offsets, patterns)) == 0


def test_kmp_vectors_from_reference_functions(oracle):
    with open(os.path.join(GOLDEN, "kmp_vectors.json")) as f:
        vectors = json.load(f)["vectors"]
    assert len(vectors) >= 500
    for v in vectors:
        pat, text = bytes.fromhex(v["pattern"]), bytes.fromhex(v["text"])
        assert oracle.kmp_prefix(pat) == v["pi"], pat
        assert oracle.kmp_count(text, pat) == v["count"], (pat, text)
        assert naive_count(text, pat) == v["count"]


def test_extract_vectors_from_reference_functions(oracle):
    with open(os.path.join(GOLDEN, "extract_vectors.json")) as f:
        vectors = json.load(f)["vectors"]
    accepted = 0
    for v in vectors:
        got = oracle.extract(bytes.fromhex(v["frame"]), v["proto"])
        if v["ok"]:
            accepted += 1
            assert got == (v["off"], v["len"]), v
        else:
            assert got is None, v
    assert accepted >= 40


def test_prefix_examples(oracle):
    # SURVEY.md 3.4, measured from the reference's kmp_prefix
    assert oracle.kmp_prefix(b"aabaaab") == [0, 1, 0, 1, 2, 2, 3]
    assert oracle.kmp_prefix(b"abcabc") == [0, 0, 0, 1, 2, 3]
    assert oracle.kmp_prefix(b"aaaa") == [0, 1, 2, 3]


def test_semantics_overlap_and_nul(oracle):
    assert oracle.kmp_count(b"aaaa", b"aa") == 3            # overlaps count (fact 3)
    assert oracle.kmp_count(b"abab\0abab", b"ab") == 2      # text stops at the first NUL (fact 1)
    assert oracle.kmp_count(b"\0abab", b"ab") == 0
    assert oracle.kmp_count(b"ab", b"abc") == 0
    assert oracle.kmp_count(b"", b"a") == 0


def test_count_csr_against_naive_on_random_batches(oracle):
    rng = random.Random(7)
    for trial in range(20):
        alpha = [b"ab", b"abc\0", bytes(range(0x20, 0x7F)) + b"\0"][trial % 3]
        pats = [bytes(rng.choice(alpha.replace(b"\0", b"")) for _ in range(rng.randint(1, 6))) for _ in range(rng.randint(0, 9))]
        pkts = [bytes(rng.choice(alpha) for _ in range(rng.choice([0, 1, 5, 40, 300]))) for _ in range(rng.randint(0, 30))]
        offsets = np.zeros(len(pkts) + 1, dtype=np.uint64)
        np.cumsum([len(p) for p in pkts], out=offsets[1:])
        data = np.frombuffer(b"".join(pkts), dtype=np.uint8)
        want = [sum(naive_count(t, p) for t in pkts) for p in pats]
        for threads in (1, 3):
            assert oracle.count_csr(data, offsets, pats, threads=threads) == want


def test_pattern_loader_edges(oracle, tmp_path):
    p = tmp_path / "s.txt"
    p.write_bytes(b"  foo\tbar\n\nfoo \x0b\x0c\r baz")
    assert oracle.load_patterns(str(p)) == [b"foo", b"bar", b"foo", b"baz"]
    p.write_bytes(b"")
    assert oracle.load_patterns(str(p)) == []
    p.write_bytes(b"x" * 99 + b" ok")
    assert oracle.load_patterns(str(p)) == [b"x" * 99, b"ok"]
    p.write_bytes(b"x" * 100)
    with pytest.raises(OSError):
        oracle.load_patterns(str(p))
    with pytest.raises(OSError):
        oracle.load_patterns(str(tmp_path / "missing.txt"))

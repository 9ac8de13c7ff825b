"""Ten-second GPU check of a library build: union-engine counts against the oracle on two bundled savefiles and 20 000
synthetic packets, then the kernel rate on 2 M packets (2.8 GB).  python scripts/quick_check.py"""
import os, sys
sys.path.insert(0, ".")
import numpy as np, torch
import multithreading_string_matching_b200 as kmp
from oracle import oracle_py
D = "tests/golden/data"
pats = kmp.load_patterns(os.path.join(D, "strings.txt"))
m = kmp.Matcher(0, engine="union"); m.set_patterns(pats)
ok = True
for f in ("udp_1000.pcap", "big_udp.pcap"):
    b = kmp.PayloadBatch(os.path.join(D, f), "udp", pinned=True)
    ok &= m.count_host(b.data, b.offsets) == oracle_py.count_csr(b.data, b.offsets, pats)
synth = kmp.Synth(seed=0xB200, payload_len=1400, plants=2, plant_patterns=pats)
sd, so = synth.fill_host(0, 20000)
ok &= m.count_host(sd, so) == oracle_py.count_csr(sd, so, pats)
n = 2_000_000
nb = synth.nbytes(0, n)
d_bytes = torch.empty(nb + 4096, dtype=torch.uint8, device="cuda:0"); d_bytes[nb:].zero_()
d_off = torch.empty(n + 1, dtype=torch.int64, device="cuda:0")
synth.fill_device(m, 0, n, d_bytes.data_ptr(), d_off.data_ptr())
d_counts = torch.zeros(len(pats), dtype=torch.int64, device="cuda:0")
st = torch.cuda.current_stream(); m.set_profile(True); ms = []
for i in range(6):
    d_counts.zero_()
    m.count_device(d_bytes.data_ptr(), d_off.data_ptr(), n, d_counts.data_ptr(), span=(0, nb), stream=st.cuda_stream)
    ms.append(m.last_kernel_ms())
print("parity_ok=%s GB/s=%.1f matches=%d" % (ok, nb / np.mean(ms[2:]) / 1e6, int(d_counts.sum().item())))

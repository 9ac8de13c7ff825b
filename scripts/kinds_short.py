"""Union-kernel rate on 2.8 GB of NUL-free random bytes (no planted tokens): the fast path almost alone.  Printed next
to every variant by scripts/gpu_variants.sh when EXTRA_PY points here."""
import sys, os, json
sys.path.insert(0, ".")
import numpy as np, torch
import multithreading_string_matching_b200 as kmp
pats = kmp.load_patterns("tests/golden/data/strings.txt")
m = kmp.Matcher(0, engine="union"); m.set_patterns(pats)
n, L = 2_000_000, 1400
off = torch.arange(0, (n + 1) * L, L, dtype=torch.int64, device="cuda:0")
g = torch.Generator(device="cuda:0"); g.manual_seed(1)
buf2 = torch.randint(1, 256, (n * L + 4096,), dtype=torch.uint8, device="cuda:0", generator=g)
d_counts = torch.zeros(len(pats), dtype=torch.int64, device="cuda:0")
st = torch.cuda.current_stream(); m.set_profile(True); ms = []
for i in range(6):
    d_counts.zero_()
    m.count_device(buf2.data_ptr(), off.data_ptr(), n, d_counts.data_ptr(), span=(0, n * L), stream=st.cuda_stream)
    ms.append(m.last_kernel_ms())
print("nonul GB/s=%.1f" % (n * L / np.mean(ms[2:]) / 1e6))

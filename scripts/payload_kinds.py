import sys, os, json
sys.path.insert(0, ".")
import numpy as np, torch
import multithreading_string_matching_b200 as kmp
pats = kmp.load_patterns("tests/golden/data/strings.txt")
m = kmp.Matcher(0, engine="union"); m.set_patterns(pats)
n, L = 2_000_000, 1400
off = torch.arange(0, (n + 1) * L, L, dtype=torch.int64, device="cuda:0")
def run(name, data):
    d_counts = torch.zeros(len(pats), dtype=torch.int64, device="cuda:0")
    st = torch.cuda.current_stream()
    m.set_profile(True); ms = []
    for i in range(5):
        d_counts.zero_()
        m.count_device(data.data_ptr(), off.data_ptr(), n, d_counts.data_ptr(), span=(0, n * L), stream=st.cuda_stream)
        ms.append(m.last_kernel_ms())
    m.set_profile(False)
    print(json.dumps({"payload": name, "GBps": n * L / np.mean(ms[2:]) / 1e6, "kernel_ms": float(np.mean(ms[2:])), "matches": int(d_counts.sum().item())}), flush=True)
g = torch.Generator(device="cuda:0"); g.manual_seed(1)
buf = torch.randint(0, 256, (n * L + 4096,), dtype=torch.uint8, device="cuda:0", generator=g)
run("uniform random bytes (a NUL every 256 bytes)", buf)
buf2 = torch.randint(1, 256, (n * L + 4096,), dtype=torch.uint8, device="cuda:0", generator=g)
run("uniform random bytes without NUL", buf2)
buf3 = torch.randint(32, 127, (n * L + 4096,), dtype=torch.uint8, device="cuda:0", generator=g)
run("printable text, no NUL, no planted tokens", buf3)
buf.zero_()
run("all NUL bytes (worst case: every group raises an event)", buf)

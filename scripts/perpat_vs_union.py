"""The two engines on the same device-resident C3 text (1 M packets x 1400 B) with the first 1 / 4 / 97 patterns of the
bundled strings.txt: kernel-only GB/s of each and whether their counts agree."""
import sys, json
sys.path.insert(0, ".")
import numpy as np, torch
import multithreading_string_matching_b200 as kmp
pats_all = kmp.load_patterns("tests/golden/data/strings.txt")
n, L = 1_000_000, 1400
synth = kmp.Synth(seed=0xB200, payload_len=L, plants=2, plant_patterns=pats_all)
for npat in (1, 4, 97):
    pats = pats_all[:npat]
    res = {}
    for eng in ("perpat", "union"):
        m = kmp.Matcher(0, engine=eng); m.set_patterns(pats)
        nbytes = synth.nbytes(0, n)
        d_bytes = torch.empty(nbytes + 4096, dtype=torch.uint8, device="cuda:0"); d_bytes[nbytes:].zero_()
        d_off = torch.empty(n + 1, dtype=torch.int64, device="cuda:0")
        synth.fill_device(m, 0, n, d_bytes.data_ptr(), d_off.data_ptr())
        d_counts = torch.zeros(len(pats), dtype=torch.int64, device="cuda:0")
        st = torch.cuda.current_stream(); m.set_profile(True); ms = []
        for i in range(4):
            d_counts.zero_()
            m.count_device(d_bytes.data_ptr(), d_off.data_ptr(), n, d_counts.data_ptr(), span=(0, nbytes), stream=st.cuda_stream)
            ms.append(m.last_kernel_ms())
        res[eng] = (nbytes / np.mean(ms[1:]) / 1e6, d_counts.cpu().tolist())
        m.close()
    print(json.dumps({"patterns": npat, "perpat_GBps": round(res["perpat"][0], 1), "union_GBps": round(res["union"][0], 1), "same_counts": res["perpat"][1] == res["union"][1]}), flush=True)

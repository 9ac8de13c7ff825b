"""Diagnostic: device-resident path vs host (chunked) path vs oracle on growing synthetic streams."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import multithreading_string_matching_b200 as kmp
from oracle import oracle_py

pats = kmp.load_patterns("tests/golden/data/strings.txt")
synth = kmp.Synth(seed=0xB200, payload_len=1400, plants=2, plant_patterns=pats)
m = kmp.Matcher(0)
m.set_patterns(pats)
for n in [int(x) for x in sys.argv[1:]] or [1000, 20000, 100000, 400000]:
    nbytes = n * 1400
    d_bytes = torch.zeros(nbytes + 4096, dtype=torch.uint8, device="cuda")
    d_off = torch.zeros(n + 1, dtype=torch.int64, device="cuda")
    synth.fill_device(m, 0, n, d_bytes.data_ptr(), d_off.data_ptr())
    d_counts = torch.zeros(len(pats), dtype=torch.int64, device="cuda")
    m.count_device(d_bytes.data_ptr(), d_off.data_ptr(), n, d_counts.data_ptr(), span=(0, nbytes))
    torch.cuda.synchronize()
    dev = d_counts.cpu().numpy()
    d_counts.zero_()
    m.count_device(d_bytes.data_ptr(), d_off.data_ptr(), n, d_counts.data_ptr(), span=(0, nbytes))
    torch.cuda.synchronize()
    dev2 = d_counts.cpu().numpy()
    hb = d_bytes[:nbytes].cpu().numpy()
    ho = d_off.cpu().numpy().astype(np.uint64)
    host = np.array(m.count_host(hb, ho))
    host2 = np.array(m.count_host(hb, ho))
    line = "n=%d dev==dev2 %s host==host2 %s dev==host %s sum dev %d host %d" % (
        n, (dev == dev2).all(), (host == host2).all(), (dev == host).all(), dev.sum(), host.sum())
    if n <= 100000:
        want = np.array(oracle_py.count_csr(hb, ho, pats))
        line += " | dev==oracle %s host==oracle %s sum %d" % ((dev == want).all(), (host == want).all(), want.sum())
        if not (dev == want).all():
            line += " diff(dev-want) " + str([(pats[i], int(dev[i] - want[i])) for i in np.nonzero(dev != want)[0][:6]])
        if not (host == want).all():
            line += " diff(host-want) " + str([(pats[i], int(host[i] - want[i])) for i in np.nonzero(host != want)[0][:6]])
    elif not (dev == host).all():
        line += " diff(dev-host) " + str([(pats[i], int(dev[i] - host[i])) for i in np.nonzero(dev != host)[0][:6]])
    print(line, flush=True)

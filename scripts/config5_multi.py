#!/usr/bin/env python3
"""BASELINE.json config 5 on N GPUs: mixed 64 / 576 / 1400 / 9000-byte payloads (40 / 20 / 30 / 10 % of the packets,
seeded), bundled strings.txt, ~1.25 GB per GPU (10 GB on 8), packets split as mpi_dumping.c:149-157 splits them.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        scripts/config5_multi.py

Two numbers, both timed as the max over ranks between barriers:
  device   every rank matches its slice, already in HBM, and the count vectors are summed with one NCCL all-reduce
  e2e      the same through kmpb_count_host: the slice waits in pinned host memory, chunked H2D on four streams
           overlapped with the kernels, counts back on the host, then the all-reduce
Rank 0 prints one JSON line."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import multithreading_string_matching_b200 as kmp  # noqa: E402
from multithreading_string_matching_b200 import distributed as kd  # noqa: E402

PER_GPU = int(os.environ.get("C5_PACKETS_PER_GPU", 855_000))
STEPS = int(os.environ.get("C5_STEPS", 10))


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    out = os.fdopen(os.dup(1), "w")  # stdout carries the JSON line only (NCCL may print to fd 1)
    os.dup2(2, 1)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    strings = kmp.load_patterns(os.path.join(ROOT, "tests", "golden", "data", "strings.txt"))
    m = kmp.Matcher(local, engine="union")
    m.set_patterns(strings)
    synth = kmp.Synth(seed=11, len_mode=1, plants=2, plant_patterns=strings)
    first, count = kd.rank_slice(PER_GPU * world, rank, world)
    nbytes = synth.nbytes(first, count)
    d_bytes = torch.zeros(nbytes + 4096, dtype=torch.uint8, device=dev)
    d_off = torch.zeros(count + 1, dtype=torch.int64, device=dev)
    synth.fill_device(m, first, count, d_bytes.data_ptr(), d_off.data_ptr())
    torch.cuda.synchronize()
    span = (int(d_off[0].item()), int(d_off[count].item()))
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t)
        return float(t.item())

    d_counts = torch.zeros(len(strings), dtype=torch.int64, device=dev)

    def device_step():
        d_counts.zero_()
        m.count_device(d_bytes.data_ptr(), d_off.data_ptr(), count, d_counts.data_ptr(), span=span, stream=stream.cuda_stream)
        kd.reduce_counts(d_counts)

    for _ in range(3):
        device_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(STEPS):
        device_step()
    e1.record()
    barrier()
    dev_ms = max_over_ranks(e0.elapsed_time(e1) / STEPS)
    resident = d_counts.cpu().tolist()

    h_bytes = torch.empty(nbytes + 4096, dtype=torch.uint8, pin_memory=True)
    h_off = torch.empty(count + 1, dtype=torch.int64, pin_memory=True)
    h_bytes[:nbytes].copy_(d_bytes[:nbytes])
    h_off.copy_(d_off - d_off[0])  # the host form takes offsets from 0
    torch.cuda.synchronize()

    def e2e_step():
        c = torch.tensor(m.count_host_ptr(h_bytes.data_ptr(), h_off.data_ptr(), count), dtype=torch.int64, device=dev)
        return kd.reduce_counts(c).cpu().tolist()

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(max(STEPS // 2, 2)):
        got = e2e_step()
    barrier()
    e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0) / max(STEPS // 2, 2))
    total_bytes = sum_over_ranks(nbytes)
    total_packets = sum_over_ranks(count)
    if rank == 0:
        print(json.dumps({
            "config": 5, "workload": "mixed 64/576/1400/9000-byte payloads (40/20/30/10 %), strings.txt", "n_gpus": world,
            "packets": int(total_packets), "payload_GB": total_bytes / 1e9,
            "device": {"GBps": total_bytes / dev_ms / 1e6, "ms": dev_ms, "packets_per_s": total_packets / dev_ms * 1e3,
                       "what": "slices resident in HBM, one NCCL all-reduce of the counts per pass"},
            "e2e": {"GBps": total_bytes / e2e_ms / 1e6, "ms": e2e_ms, "packets_per_s": total_packets / e2e_ms * 1e3,
                    "what": "kmpb_count_host from pinned host memory (H2D on four streams overlapped with the kernels) + all-reduce"},
            "host_and_device_paths_agree": got == resident, "matches": int(sum(resident))}), file=out, flush=True)
    m.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

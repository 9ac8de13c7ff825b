"""N > 1 host logic on CPU: two gloo ranks, packets split by the mpi_dumping.c rule, counts summed by
all-reduce.  The per-slice counting is done by the oracle here (there is no GPU); on the B200 box the
same driver runs with the CUDA matcher (tests/test_gpu_parity.py::test_sharded_slices_add_up)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import multithreading_string_matching_b200 as kmp
from multithreading_string_matching_b200 import distributed as kd

from conftest import DATA

N_PACKETS = 1001  # odd on purpose: rank 0 takes the remainder


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import oracle_py

    patterns = kmp.load_patterns(os.path.join(DATA, "strings.txt"))
    synth = kmp.Synth(seed=0xB200, payload_len=300, plants=2, plant_patterns=patterns)

    def count_slice(first, count):
        data, off = synth.fill_host(first, count)  # every rank generates only its own slice
        return oracle_py.count_csr(data, off, patterns, threads=1)

    first, count = kd.rank_slice(N_PACKETS)
    total = kd.sharded_count(count_slice, N_PACKETS, len(patterns))
    torch.save({"first": first, "count": count, "total": total}, os.path.join(out, "rank%d.pt" % rank))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_split_and_reduce(tmp_path, oracle):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    res = [torch.load(os.path.join(str(tmp_path), "rank%d.pt" % r)) for r in range(world)]
    assert (res[0]["first"], res[0]["count"]) == (0, 501)  # N/P + N%P
    assert (res[1]["first"], res[1]["count"]) == (501, 500)
    patterns = kmp.load_patterns(os.path.join(DATA, "strings.txt"))
    data, off = kmp.Synth(seed=0xB200, payload_len=300, plants=2, plant_patterns=patterns).fill_host(0, N_PACKETS)
    want = oracle.count_csr(data, off, patterns)
    assert res[0]["total"] == want and res[1]["total"] == want and sum(want) > 0


def test_single_process_is_identity():
    assert kd.world() == (0, 1)
    assert kd.rank_slice(10) == (0, 10)
    t = torch.tensor([1, 2, 3])
    assert kd.reduce_counts(t).tolist() == [1, 2, 3]

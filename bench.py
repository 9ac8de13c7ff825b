#!/usr/bin/env python3
"""bench.py -- payload GB/s of the KMP packet-matching hot path on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2]): synthetic UDP stream, 10 M packets x 1400-byte payloads per GPU
(counter-based generator, csrc/cuda/synth.cu; last payload byte NUL, 2 strings.txt tokens planted per
packet), matched against the 97 patterns of the bundled strings.txt.  Packets are split over ranks as
mpi_dumping.c:149-157 splits them; weak scaling: the stream grows with N, each GPU matches its own
contiguous slice and the 97-entry count vectors are summed with one NCCL all-reduce (mpi_dumping.c:202).

A "step" is one pass of the hot path over the rank's whole slice.
  value  device-resident: payload already in HBM when the timed region starts (CUDA events, max over ranks)
  e2e    the same pass through the C ABI's host entry point kmpb_count_host: pinned host CSR -> chunked
         H2D on 4 streams overlapped with the kernels -> counts back on the host
  roofline      the union kernel alone (events around the kernel on its stream) vs the measured HBM peak
  cpu_baseline  the unmodified reference on a bounded prefix of the same stream: oracle/_ref/openmp_data on all host
                threads, and under "serial" oracle/_ref/serial on one core (a shorter prefix) -- reported baselines,
                not the target
--impl reference times that CPU program as the step itself.
"""
import argparse
import ctypes
import json
import os
import resource
import shutil
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
DATA = os.path.join(ROOT, "tests", "golden", "data")
REFBIN = os.path.join(ROOT, "oracle", "_ref")
SEED = 0xB200


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md; MEASURED_PEAKS.json absent)"


class ClockSampler:
    """SM clock and throttle reasons sampled every few ms WHILE the timed region runs: NVML in a thread
    (nvidia_ml_py), nvidia-smi -lms as the fallback."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = self.path = self.thread = None
        self.samples, self.reasons, self.sm_max = [], set(), None
        self.stop_flag = threading.Event()

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(self.gpu).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
        except Exception:
            visible = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v for v in visible.split(",") if v.strip().isdigit()]
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(int(ids[self.gpu]) if self.gpu < len(ids) else self.gpu)

    def _loop(self, pynvml, h):
        while not self.stop_flag.is_set():
            try:
                self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                bits = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for name, bit in self.REASONS:
                    if bits & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        try:
            pynvml, h = self._nvml_handle()
            self.sm_max = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._loop, args=(pynvml, h), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.thread = None
        if shutil.which("nvidia-smi") is None:
            return
        fd, self.path = tempfile.mkstemp(suffix=".csv")
        os.close(fd)
        self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS,
                                      "--format=csv,noheader,nounits", "-lms", "20"],
                                     stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "how": None}
        if self.thread is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            if self.samples:
                out.update(sm_mhz=float(np.median(self.samples)), sm_max_mhz=self.sm_max, reasons=sorted(self.reasons),
                           samples=len(self.samples), how="NVML in a thread during the timed steps")
            return out
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, reasons, smax = [], set(), None
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=smax, reasons=sorted(reasons), samples=len(sm),
                       how="nvidia-smi -lms 20 during the timed steps")
        return out


def measured_traffic(packets, payload_len):
    """DRAM bytes per launch of the dominant kernel from the committed ncu --set full capture of this
    workload (profiles/r01_traffic.json), or None when the workload differs."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if int(t["packets"]) == int(packets) and int(t["payload_len"]) == int(payload_len):
            return int(t["dram_bytes_read"]) + int(t["dram_bytes_write"])
    except Exception:
        pass
    return None


# ---- CPU reference arm ---------------------------------------------------------------------------

def write_pcap(path, data, offsets):
    """Classic LE pcap v2.4, linktype 1, caplen == len; frame = Ethernet(0x0800) + IPv4(0x45, proto 17)
    + UDP + payload."""
    n = len(offsets) - 1
    lens = np.diff(offsets.astype(np.int64))
    with open(path, "wb") as f:
        f.write(np.array([0xA1B2C3D4, 0x00040002, 0, 0, 262144, 1], dtype="<u4").tobytes())
        if n and (lens == lens[0]).all():
            L = int(lens[0])
            rec = np.zeros((n, 16 + 42 + L), dtype=np.uint8)
            hdr = np.zeros(58, dtype=np.uint8)
            hdr[8:12] = np.frombuffer(np.uint32(42 + L).tobytes(), dtype=np.uint8)
            hdr[12:16] = hdr[8:12]
            hdr[16 + 12:16 + 14] = (0x08, 0x00)
            hdr[16 + 14] = 0x45
            hdr[16 + 16:16 + 18] = ((28 + L) >> 8, (28 + L) & 255)
            hdr[16 + 22], hdr[16 + 23] = 64, 17
            hdr[16 + 38:16 + 40] = ((8 + L) >> 8, (8 + L) & 255)
            rec[:, :58] = hdr
            rec[:, 58:] = data[: n * L].reshape(n, L)
            rec.tofile(f)
        else:
            for k in range(n):
                L = int(lens[k])
                hdr = bytearray(58)
                hdr[8:12] = hdr[12:16] = int(42 + L).to_bytes(4, "little")
                hdr[28:30] = b"\x08\x00"
                hdr[30] = 0x45
                hdr[32:34] = int(28 + L).to_bytes(2, "big")
                hdr[38], hdr[39] = 64, 17
                hdr[54:56] = int(8 + L).to_bytes(2, "big")
                f.write(hdr)
                f.write(data[int(offsets[k]):int(offsets[k + 1])].tobytes())


def run_reference_program(pcap, strings, threads, program="openmp_data"):
    """oracle/_ref/openmp_data or oracle/_ref/serial (the unmodified reference, -O2) -> (self-reported seconds,
    wall seconds, stdout).  serial.c takes no thread count and times reading the savefile too (serial.c:110-160)."""
    exe = os.path.join(REFBIN, program)

    def unlimited_stack():  # openmp_data.c:123 puts one pointer per packet on the stack
        try:
            resource.setrlimit(resource.RLIMIT_STACK, (resource.RLIM_INFINITY, resource.RLIM_INFINITY))
        except Exception:
            pass

    env = {k: v for k, v in os.environ.items() if not k.startswith("MALLOC_")}
    t0 = time.perf_counter()
    argv = [exe, pcap, strings] + ([str(threads)] if program == "openmp_data" else [])
    out = subprocess.run(argv, capture_output=True, env=env, preexec_fn=unlimited_stack, check=True).stdout
    wall = time.perf_counter() - t0
    lines = out.decode("latin-1").splitlines()
    return float(lines[-1].split("=")[1].split()[0]), wall, "\n".join(lines[:-1]) + "\n"


class CpuReference:
    """The CPU arm: reference binary if it was built (kind 'reference'), else the oracle port."""

    def __init__(self, kmp, patterns, payload_len):
        self.kmp, self.patterns, self.payload_len = kmp, patterns, payload_len
        self.cores = os.cpu_count() or 1
        self.kind = "reference" if os.path.isfile(os.path.join(REFBIN, "openmp_data")) else "port"
        self.tmp = tempfile.mkdtemp(prefix="kmpb_ref_")
        self.synth = kmp.Synth(seed=SEED, payload_len=payload_len, plants=2, plant_patterns=patterns)

    def prepare(self, n_packets):
        self.n = n_packets
        self.data, self.offsets = self.synth.fill_host(0, n_packets)
        if self.kind == "reference":
            self.pcap = os.path.join(self.tmp, "sample.pcap")
            write_pcap(self.pcap, self.data, self.offsets)

    def run(self):
        """-> (seconds of the path, counts as reported)."""
        if self.kind == "reference":
            secs, wall, text = run_reference_program(self.pcap, os.path.join(DATA, "strings.txt"), self.cores)
            return secs, text
        from oracle import oracle_py
        t0 = time.perf_counter()
        counts = oracle_py.count_csr(self.data, self.offsets, self.patterns, threads=self.cores)
        return time.perf_counter() - t0, oracle_py.format_report(self.patterns, counts).decode("latin-1")

    def calibrate(self, target_seconds=12.0):
        """Pick a sample size that takes about target_seconds on this box's cores."""
        self.prepare(4000)
        secs, _ = self.run()
        rate = 4000 / max(secs, 1e-3)
        n = int(min(max(rate * target_seconds, 4000), 400_000))
        self.prepare(n)
        return n

    def sample_text(self):
        return "first %d packets x %d B of the bench stream (seed 0x%X), %s" % (
            self.n, self.payload_len, SEED,
            "oracle/_ref/openmp_data -O2 self-reported Elapsed time" if self.kind == "reference" else "oracle port (OpenMP)")

    def close(self):
        shutil.rmtree(self.tmp, ignore_errors=True)


def reference_arm(args, kmp, patterns):
    rank = env_int("RANK", 0)
    if rank != 0:
        return
    ref = CpuReference(kmp, patterns, args.payload_len)
    n = args.ref_packets or ref.calibrate(8.0)
    if args.ref_packets:
        ref.prepare(n)
    times = []
    for i in range(args.warmup + args.steps):
        secs, _ = ref.run()
        if i >= args.warmup:
            times.append(secs)
    ref.close()
    payload = n * args.payload_len
    ms = 1e3 * sum(times) / len(times)
    value = payload / (ms / 1e3) / 1e9
    line = {
        "impl": "reference", "metric": "payload_GBps", "value": value, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "synthetic UDP pcap, %d packets x %d B payloads per GPU, bundled strings.txt (97 patterns, 87 distinct)"
                               % (args.packets, args.payload_len),
                   "sample": "each step = the first %d packets of that stream on the host cores" % n,
                   "packets_per_step": n, "payload_bytes_per_step": payload},
        "packets_per_s": n / (ms / 1e3),
        "cpu_baseline": {"value": value, "unit": "GB/s", "cores": ref.cores, "kind": ref.kind, "sample": ref.sample_text()},
        "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---- our arm --------------------------------------------------------------------------------------

def ours(args, kmp, patterns):
    import torch
    import torch.distributed as dist

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)

    per_gpu = args.packets
    total_packets = per_gpu * world                      # weak scaling: the stream grows with N
    from multithreading_string_matching_b200 import distributed as kd
    first, count = kd.rank_slice(total_packets, rank, world)   # mpi_dumping.c:149-157
    L = args.payload_len
    n_pat = len(patterns)
    m = kmp.Matcher(local, engine=args.engine)
    m.set_patterns(patterns)
    synth = kmp.Synth(seed=SEED, payload_len=L, plants=2, plant_patterns=patterns)
    nbytes = synth.nbytes(first, count)
    d_bytes = torch.empty(nbytes + 4096, dtype=torch.uint8, device=dev)
    d_bytes[nbytes:].zero_()
    d_off = torch.empty(count + 1, dtype=torch.int64, device=dev)
    synth.fill_device(m, first, count, d_bytes.data_ptr(), d_off.data_ptr())
    # two count vectors: the all-reduce of step i runs on a side stream while step i+1's kernel matches the
    # next batch (in a pipeline the batches differ; here it is the same slice again)
    d_counts2 = [torch.zeros(n_pat, dtype=torch.int64, device=dev) for _ in range(2)]
    d_counts = d_counts2[0]
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream()
    side = torch.cuda.Stream(device=dev) if world > 1 else None
    reduced = [None, None]
    step_no = [0]

    # Reduce across GPUs.  Preferred: every rank's count vectors live in symmetric memory (mapped into all
    # ranks over NVLink) and the match kernel's last block adds its counts straight into all of them --
    # no collective call at all (kmpb_count_device_span_peers).  One row of counts per step, so no rank
    # ever clears a vector another rank may be adding to.  Fallback: one NCCL all-reduce per step on a
    # side stream, overlapped with the next step's kernel.
    p2p = None
    if world > 1 and args.reduce != "nccl" and args.engine == "union":
        ok, why = 1, ""
        try:
            import torch.distributed._symmetric_memory as symm
            rows = args.warmup + args.steps + 16
            sym = symm.empty((rows, n_pat), dtype=torch.int64, device=dev)
            sym.zero_()
            hdl = symm.rendezvous(sym, dist.group.WORLD)
            peer_ptrs = [int(hdl.buffer_ptrs[r]) for r in range(world)]
        except Exception as e:  # no symmetric memory on this box / build
            ok, why = 0, repr(e)[:200]
        agree = torch.tensor([ok], dtype=torch.int32, device=dev)
        dist.all_reduce(agree, op=dist.ReduceOp.MIN)
        if int(agree.item()) == 1:
            p2p = {"sym": sym, "ptrs": peer_ptrs, "rows": rows}
        elif rank == 0 and why:
            print("bench: symmetric memory unavailable, using NCCL all-reduce: " + why, file=sys.stderr)
        torch.cuda.synchronize()
        dist.barrier()

    def step():
        if p2p is not None:
            row = step_no[0] % p2p["rows"]
            step_no[0] += 1
            vecs = [ptr + row * n_pat * 8 for ptr in p2p["ptrs"]]
            m.count_device_into(d_bytes.data_ptr(), d_off.data_ptr(), count, vecs, span=(0, nbytes), stream=stream.cuda_stream)
            return p2p["sym"][row]
        i = step_no[0] & 1
        step_no[0] += 1
        buf = d_counts2[i]
        if reduced[i] is not None:
            stream.wait_event(reduced[i])  # the vector's previous all-reduce has read it
        buf.zero_()
        m.count_device(d_bytes.data_ptr(), d_off.data_ptr(), count, buf.data_ptr(), span=(0, nbytes),
                       stream=stream.cuda_stream)
        if world > 1:
            ready = torch.cuda.Event()
            ready.record(stream)
            with torch.cuda.stream(side):
                side.wait_event(ready)
                kd.reduce_counts(buf)  # the MPI_Reduce(SUM) of mpi_dumping.c:202, as an NCCL all-reduce over NVLink
                reduced[i] = torch.cuda.Event()
                reduced[i].record(side)
        return buf

    def drain_side():
        if world > 1 and p2p is None:
            stream.wait_stream(side)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing -------------------------------------------------------------------
    for _ in range(args.warmup):
        step()
    drain_side()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = m.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        d_counts = step()
    drain_side()  # the last all-reduce is inside the timed region
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms_step = max_over_ranks(e0.elapsed_time(e1)) / args.steps
    launches = m.launches - launches0  # partition + union + count expansion per step (libkmpb200's own counter)
    counts_resident = d_counts.cpu().numpy().copy()

    # ---- the dominant kernel alone (roofline) -----------------------------------------------------
    m.set_profile(True)
    kms = []
    for _ in range(max(3, min(args.steps, 10))):
        step()
        kms.append(m.last_kernel_ms())
    drain_side()
    m.set_profile(False)
    kernel_ms = max_over_ranks(float(np.mean(kms)))
    algo_bytes = nbytes + 8 * (count + 1)  # payload once + one offset per packet
    peak, peak_src = measured_hbm_peak()
    achieved = algo_bytes / (kernel_ms / 1e3) / 1e9

    # ---- end to end through the host entry point -------------------------------------------------
    e2e = None
    if not args.no_e2e:
        import psutil
        need = (nbytes + 8 * (count + 1)) * world
        e2e_count = count
        if psutil.virtual_memory().available < 2.5 * need:
            e2e_count = max(int(count * psutil.virtual_memory().available / (3.0 * need)), 1000)
        e_bytes = synth.nbytes(first, e2e_count)
        h_bytes = torch.empty(e_bytes + 4096, dtype=torch.uint8, pin_memory=True)
        h_off = torch.empty(e2e_count + 1, dtype=torch.int64, pin_memory=True)
        h_bytes[: e_bytes].copy_(d_bytes[: e_bytes])
        h_off.copy_(d_off[: e2e_count + 1])
        torch.cuda.synchronize()
        e2e_steps = max(1, min(args.steps, 5))

        def e2e_step():
            c = m.count_host_ptr(h_bytes.data_ptr(), h_off.data_ptr(), e2e_count)
            if world > 1:
                t = torch.tensor(c, dtype=torch.int64, device=dev)
                dist.all_reduce(t)
                c = t.cpu().tolist()
            return c

        c = e2e_step()
        barrier()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(e2e_steps):
            c = e2e_step()
        e1.record()
        barrier()
        wall_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps
        e2e_ms = max_over_ranks(max(e0.elapsed_time(e1) / e2e_steps, wall_ms))
        if e2e_count == count:
            assert np.array_equal(np.asarray(c, dtype=np.int64), counts_resident), "host and device paths disagree"
        e2e_total = e_bytes * world if world == 1 else None
        if world > 1:
            t = torch.tensor([float(e_bytes)], dtype=torch.float64, device=dev)
            dist.all_reduce(t)
            e2e_total = float(t.item())
        e2e = {"value": e2e_total / (e2e_ms / 1e3) / 1e9, "unit": "GB/s", "ms_per_step": e2e_ms,
               "h2d_bytes_per_step": int(e_bytes + 8 * (e2e_count + 1)), "d2h_bytes_per_step": int(8 * n_pat),
               "packets_per_step_per_gpu": int(e2e_count), "steps": e2e_steps,
               "note": "kmpb_count_host: pinned host CSR -> 64 MiB chunks over 4 streams (H2D overlapped with kernels) -> counts D2H"
                       + ("" if e2e_count == count else "; REDUCED sample: host RAM too small for the full slice")}
        del h_bytes, h_off

    # total payload over all ranks
    total_bytes = float(nbytes)
    if world > 1:
        t = torch.tensor([total_bytes], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        total_bytes = float(t.item())

    if rank == 0:
        value = total_bytes / (ms_step / 1e3) / 1e9
        line = {
            "metric": "payload_GBps", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "synthetic UDP pcap, %d packets x %d B payloads per GPU, bundled strings.txt (97 patterns, 87 distinct)"
                                   % (per_gpu, L),
                       "packets_total": total_packets, "payload_bytes_total": int(total_bytes), "engine": args.engine,
                       "split": "mpi_dumping.c:149-157 contiguous packet slices",
                       "reduce": ("none (one GPU)" if world == 1 else
                                  "in the match kernel: its last block adds the counts to every rank's vector in symmetric memory over NVLink (no collective call)"
                                  if p2p is not None else "one NCCL all-reduce per step on a side stream, overlapped with the next step's kernel"),
                       "l2": "inputs (%.1f GB per GPU) larger than the 126 MB L2; no flush needed" % (nbytes / 1e9)},
            "packets_per_s": total_packets / (ms_step / 1e3),
            "hbm_frac_of_measured_peak": value / world / peak,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(count, L) if args.engine != "perpat" else None, "kernel": "kmpb_union_kernel" if args.engine != "perpat" else "kmpb_perpat_kernel",
                         "kernel_ms": kernel_ms, "algorithmic_bytes_per_launch": int(algo_bytes), "peak_source": peak_src},
            "e2e": e2e, "gpu_launches": int(launches), "gpu_launches_per_step": int(launches) // max(args.steps, 1), "clocks": clocks,
            "matches_per_step": int(counts_resident.sum()),
        }
        if world == 1 and not args.no_cpu:
            ref = CpuReference(kmp, patterns, L)
            n = args.ref_packets or ref.calibrate(12.0)
            if args.ref_packets:
                ref.prepare(n)
            secs, text = ref.run()
            # parity at bench time: our counts on the same prefix must print the same lines
            d_c = torch.zeros(n_pat, dtype=torch.int64, device=dev)
            m.count_device(d_bytes.data_ptr(), d_off.data_ptr(), n, d_c.data_ptr(), span=(0, n * L), stream=stream.cuda_stream)
            torch.cuda.synchronize()
            ours_text = kmp.format_report(patterns, d_c.cpu().tolist()).decode("latin-1")
            line["cpu_baseline"] = {"value": n * L / secs / 1e9, "unit": "GB/s", "cores": ref.cores, "kind": ref.kind,
                                    "sample": ref.sample_text(), "seconds": secs, "counts_match_gpu": ours_text == text}
            if ref.kind == "reference" and os.path.isfile(os.path.join(REFBIN, "serial")):
                # serial.c, the one-core form of the same loop, on a prefix of that sample sized for a few seconds
                ns = max(min(n // (2 * ref.cores), n), 1000)
                spcap = os.path.join(ref.tmp, "serial.pcap")
                write_pcap(spcap, ref.data[: ns * L], ref.offsets[: ns + 1])
                ssecs, _, stext = run_reference_program(spcap, os.path.join(DATA, "strings.txt"), 1, program="serial")
                d_c.zero_()
                m.count_device(d_bytes.data_ptr(), d_off.data_ptr(), ns, d_c.data_ptr(), span=(0, ns * L), stream=stream.cuda_stream)
                torch.cuda.synchronize()
                line["cpu_baseline"]["serial"] = {
                    "value": ns * L / ssecs / 1e9, "unit": "GB/s", "cores": 1, "seconds": ssecs,
                    "sample": "first %d packets of the same stream, oracle/_ref/serial -O2 self-reported Elapsed time (includes reading the savefile)" % ns,
                    "counts_match_gpu": kmp.format_report(patterns, d_c.cpu().tolist()).decode("latin-1") == stext}
            if ref.kind == "reference":
                # the step before the path: savefile -> pinned CSR batch, on the same sample pcap
                t0 = time.perf_counter()
                batch = kmp.PayloadBatch(ref.pcap, "udp", pinned=True)
                dt = time.perf_counter() - t0
                line["ingest"] = {"value": batch.total_bytes / dt / 1e9, "unit": "GB/s", "seconds": dt, "packets": int(batch.n_packets),
                                  "what": "kmpb_load_pcap_csr (mmap, sequential record framing, OpenMP pack into pinned memory) on the CPU sample's pcap"}
                batch.close()
            ref.close()
        print(json.dumps(line), flush=True)
    m.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--packets", type=int, default=10_000_000, help="packets per GPU (BASELINE config 3: 10 M)")
    ap.add_argument("--payload-len", type=int, default=1400)
    ap.add_argument("--engine", default="union", choices=["union", "perpat"])
    ap.add_argument("--ref-packets", type=int, default=0, help="CPU sample size (0 = calibrate to ~10 s)")
    ap.add_argument("--reduce", default="auto", choices=["auto", "nccl"], help="N>1: in-kernel reduce over symmetric memory, or NCCL")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    # stdout carries the one JSON line and nothing else: whatever libraries write to file descriptor 1 in the
    # meantime (NCCL prints its version there when the box sets NCCL_DEBUG) goes to stderr
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = json_out

    import multithreading_string_matching_b200 as kmp  # raises if libkmpb200.so is missing: no fallback

    patterns = kmp.load_patterns(os.path.join(DATA, "strings.txt"))
    if args.impl == "reference":
        reference_arm(args, kmp, patterns)
    else:
        ours(args, kmp, patterns)


if __name__ == "__main__":
    main()

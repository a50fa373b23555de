#!/usr/bin/env python
"""bench.py — headline benchmark of the SA + BWT (and inverse BWT) hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--mode sharded|independent]

One "step" = one pass of the hot path over one synthetic text: suffix array (n+1 int32) AND forward
BWT (n bytes + sentinel index) of the text.
  * value  : input MB/s, text already resident in HBM, results left in HBM (one sort + one BWT pass per
             step through b200sa_bwt_dev), CUDA events on the launching stream, max over ranks.
  * e2e    : the same metric through the reference-facing host entry points with HOST buffers — the
             two drop-in calls a user of the reference makes (make_suffix_array then
             forward_burrows_wheeler_transform => b200sa_suffix_array + b200sa_bwt; the second call recognises the
             resident text and reuses the sort), pinned host memory, H2D and D2H copies inside the timed region.
  * e2e_facade: the same two calls through the C++ facade with pageable std::vector storage (tools/facade_bench.cpp).
  * unbwt  : the inverse transform of the 2^30-2 byte Markov text (BASELINE.json configs[3]): device-resident value,
             e2e through b200sa_unbwt, roofline of the walk, the reference's inverse beside it.
  * roofline: the dominant kernel (k_onesweep_pass, radix scatter sweeps): algorithmic bytes (24 B per
             tuple per sweep, 20 B for the first sweep of round 0 whose values are generated) / summed
             CUDA-event time of those launches inside the timed region / measured HBM copy peak.
  * cpu_baseline: the UNMODIFIED reference (oracle/_ref, built from /root/reference) timed on this
             box's host cores on the same text.
N > 1 (default --mode sharded): ONE text of the same configuration sharded over the N ranks — key-range partition of
the suffixes, the inverse suffix array in NVLink peer memory, round loop in C++ (b200sa_shard_sort); strong scaling,
value = n / max-over-ranks time.  The line also carries the 2^30-2 ACGT text (configs[2]) and the inverse BWT of the
2^30-2 Markov text (configs[3]) on the same N GPUs.  --mode independent: one text per GPU, no exchange (weak scaling).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (generator, n, description)
    "markov3_256MiB": ("markov3", 1 << 28, "256 MiB synthetic order-3 Markov English-like text: SA + BWT (BASELINE.json configs[1])"),
    "rand_16MiB": ("rand", 1 << 24, "16 MiB synthetic random bytes: SA + BWT (BASELINE.json configs[0])"),
    "markov3_64MiB": ("markov3", 1 << 26, "64 MiB Markov text (reduced; debugging only)"),
    "acgt_1GiB": ("acgt_rep", (1 << 30) - 2, "2^30-2 ACGT bases with injected repeats: SA + BWT (BASELINE.json configs[2])"),
    "markov3_1GiB": ("markov3", (1 << 30) - 2, "2^30-2 bytes of the Markov text (BASELINE.json configs[3]: its BWT is the inverse transform's input)"),
}
UNBWT_WORKLOAD = "markov3_1GiB"
CPU_MAX_BYTES = 1 << 28     # reference arm / cpu_baseline: the whole text up to 256 MiB (the reference needs ~8 s per step there)
REF_BUILD_NOTE = "oracle/_ref: unmodified reference, g++ -O3 -march=x86-64-v3 (portable across build and GPU box; the reference's own flags use -march=native)"


def cpu_model() -> str:
    """the host CPU the reference arm ran on (print the core count and CPU model beside the CPU baseline)"""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.lower().startswith("model name"):
                    return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


def gen_text(kind: str, n: int, seed_offset: int = 0) -> np.ndarray:
    from msufsort_b200 import textgen as t
    if kind == "markov3":
        return t.markov3(n, t.SEED_MARKOV + seed_offset)
    if kind == "rand":
        return t.rand(n, t.SEED_RAND + seed_offset)
    if kind == "acgt_rep":
        return t.acgt_rep(n, t.SEED_ACGT + seed_offset)
    raise ValueError(kind)


# ---------------------------------------------------------------------------------------------
# reference arm: the unmodified reference library on host cores (oracle/_ref)

def load_reference():
    path = os.path.join(ROOT, "oracle", "_ref", "libmsufsort_ref.so")
    if not os.path.exists(path):
        return None
    lib = C.CDLL(path)
    P, I64, I32 = C.c_void_p, C.c_int64, C.c_int32
    lib.ref_make_suffix_array.argtypes = [P, I64, P, I32]
    lib.ref_forward_bwt.argtypes = [P, I64, I32]
    lib.ref_forward_bwt.restype = I32
    lib.ref_reverse_bwt.argtypes = [P, I64, I32, I32]
    lib.ref_hardware_concurrency.restype = C.c_int
    return lib


def reference_step(lib, text: np.ndarray, sa_out: np.ndarray, work: np.ndarray, threads: int):
    """one step on the CPU: make_suffix_array + forward_burrows_wheeler_transform (its two public calls);
    returns (seconds, sentinel index); `work` holds the BWT afterwards"""
    np.copyto(work, text)
    t0 = time.perf_counter()
    lib.ref_make_suffix_array(text.ctypes.data, text.size, sa_out.ctypes.data, threads)
    s = lib.ref_forward_bwt(work.ctypes.data, work.size, threads)
    return time.perf_counter() - t0, int(s)


def reference_unbwt(lib, bwt: np.ndarray, sentinel: int, threads: int, want: np.ndarray):
    """the reference's inverse transform of `bwt` in place; returns seconds (and checks the round trip)"""
    t0 = time.perf_counter()
    lib.ref_reverse_bwt(bwt.ctypes.data, bwt.size, sentinel, threads)
    dt = time.perf_counter() - t0
    if not np.array_equal(bwt, want):
        raise SystemExit("bench.py: the reference's inverse BWT did not restore the text")
    return dt


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    kind, n, desc = WORKLOADS[wl]
    lib = load_reference()
    if lib is None:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libmsufsort_ref.so not built (reference tree absent at build time)"}))
        return 0
    threads = lib.ref_hardware_concurrency()
    ns = min(n, CPU_MAX_BYTES)
    text = gen_text(kind, ns)
    sa = np.empty(ns + 1, dtype=np.int32)
    work = np.empty(ns, dtype=np.uint8)
    for _ in range(args.warmup):
        reference_step(lib, text, sa, work, threads)
    times, sentinel = [], 0
    for _ in range(args.steps):
        dt, sentinel = reference_step(lib, text, sa, work, threads)
        times.append(dt)
    total = sum(times)
    value = ns * args.steps / total / 1e6
    # the inverse transform of the BWT the last step left in `work` (one repetition; the reference needs ~4 s at 256 MiB)
    un_s = reference_unbwt(lib, work, sentinel, threads, text)
    # the same step with numThreads = 1 beside it (one repetition, ~16 s at 256 MiB; N = 1 runs only)
    single = None
    if args.gpus == 1 and not args.no_single_thread:
        t1, s1 = reference_step(lib, text, sa, work, 1)
        single = {"value": ns / t1 / 1e6, "unit": "MB/s", "ms_per_step": 1e3 * t1, "cores": 1, "sample": "one repetition of the same step, numThreads = 1"}
        assert s1 == sentinel
    whole = "the whole workload text" if ns == n else f"the first {ns} bytes of the workload text"
    sample = f"{whole}, SA + BWT via the reference's two public calls, {threads} threads; {REF_BUILD_NOTE}"
    line = {
        "impl": "reference", "metric": "sa_bwt_input_throughput", "value": value, "unit": "MB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.gpus > 1 and args.mode == "sharded" else "weak", "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
        "config": {"workload": wl, "description": desc, "n_bytes": n, "sample_bytes": ns},
        "cpu_baseline": {"value": value, "unit": "MB/s", "cores": threads, "cpu_model": cpu_model(), "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "MB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "unbwt": {"value": ns / un_s / 1e6, "unit": "MB/s", "ms": 1e3 * un_s, "cores": threads,
                  "sample": f"reverse_burrows_wheeler_transform of the BWT of {whole} ({ns} bytes), 1 repetition"},
        "single_thread": single,
    }
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------
# our arm

def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    from msufsort_b200.api import torch_stream_handle, Engine

    kind, n, desc = WORKLOADS[wl]
    if args.n:
        n = args.n
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this framework has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    eng = Engine(local_rank)
    stream = torch_stream_handle()

    # every rank gets its own text of the same size (rank 0 = the seed the parity tests use)
    host_text = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    np.copyto(host_text.numpy(), gen_text(kind, n, seed_offset=rank * 7919))
    d_text = host_text.cuda(non_blocking=False)
    d_sa = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    d_bwt = torch.empty(n, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident: value + roofline
    # the clock sampler starts before the warm-up steps (nvidia-smi needs ~100 ms to come up) and
    # stops right after the timed region: every sample is taken under the same load
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        eng.bwt_dev(d_text, n, d_bwt, d_sa, stream)
    barrier()
    eng.profile_reset()
    eng.set_profiling(True)
    launches0 = eng.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        sentinel = eng.bwt_dev(d_text, n, d_bwt, d_sa, stream)
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = eng.launch_count() - launches0
    prof = eng.profile()
    eng.set_profiling(False)
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        tt = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_ms = float(tt.item())

    # quick sanity inside the bench (outside the timed region): the GPU validator must accept the SA
    bad = eng.check_suffix_array_dev(d_text, n, d_sa, stream)
    if bad != 0:
        raise SystemExit(f"bench.py: validator found {bad} bad rows — refusing to report a number")

    # ---- end to end through the host entry points (pinned buffers, copies inside the timed region)
    h_sa = torch.empty(n + 1, dtype=torch.int32, pin_memory=True)
    h_work = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    e2e_times = []
    e2e_launches = 0
    for it in range(min(args.warmup, 2) + args.steps):
        h_work.copy_(host_text)  # in-place API: restore the caller's buffer outside the timed region
        barrier()
        l0 = eng.launch_count()
        t0 = time.perf_counter()
        eng.suffix_array_ptr(host_text.data_ptr(), n, h_sa.data_ptr())
        s2 = eng.bwt_ptr(h_work.data_ptr(), n)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if it >= min(args.warmup, 2):
            e2e_times.append(dt)
            e2e_launches = eng.launch_count() - l0
    e2e_s = sum(e2e_times) / len(e2e_times)
    if world > 1:
        tt = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    assert s2 == sentinel
    assert int(h_sa[0]) == n and int(h_sa[sentinel]) == 0
    assert bool(torch.equal(h_work.cuda(), d_bwt)), "the end-to-end BWT differs from the device-resident one"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = load_peaks()
    sp = prof["phases"]["sort_pass"]
    achieved = sp["alg_bytes"] / (sp["ms"] * 1e-3) / 1e9 if sp["ms"] > 0 else 0.0
    # DRAM traffic per launch: NOT measured in this run (that needs a profiler).  It is the DRAM / algorithmic byte ratio of the
    # committed ncu --set full capture of this kernel (profiles/*_traffic.json) applied to this run's algorithmic bytes per launch.
    traffic, traffic_src = None, None
    for name in ("r02_traffic.json", "r01_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                tr = json.load(f)
            traffic = tr["dram_bytes_per_launch"] / tr["algorithmic_bytes_per_launch"] * sp["alg_bytes"] / sp["launches"]
            traffic_src = f"from profiles/{name} (ncu --set full of the m=2^28 sweep: dram__bytes_read.sum + dram__bytes_write.sum = {tr['dram_bytes_per_launch']:.4g} B " \
                          f"for {tr['algorithmic_bytes_per_launch']:.4g} algorithmic bytes), scaled to this run's algorithmic bytes per launch; not measured in this run"
            break
        except Exception:
            continue
    ms_per_step = dev_ms / args.steps
    value = world * n / (ms_per_step * 1e-3) / 1e6
    phases = {k: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps,
                  "alg_GB_per_step": v["alg_bytes"] / args.steps / 1e9} for k, v in prof["phases"].items() if v["launches"]}

    # ---- the rows around the hot path, measured after and outside the timed regions above: LCP array of
    # the same text from the finished SA, and the text cut into 1024 blocks transformed as ONE batch (forward + inverse)
    extras = None
    if world == 1 and not args.no_extras:
        try:
            extras = measure_extras(eng, torch, d_text, d_sa, d_bwt, n, stream)
        except Exception as exc:  # the headline line must not depend on the extras
            extras = {"error": str(exc)[:200]}

    # ---- CPU baseline: the reference itself on the same text (rank 0, N=1 only): SA + BWT, then its inverse
    cpu, cpu_unbwt = None, None
    if world == 1 and not args.no_cpu_baseline:
        lib = load_reference()
        if lib is not None:
            threads = lib.ref_hardware_concurrency()
            ns = min(n, CPU_MAX_BYTES)
            text = host_text.numpy()[:ns].copy()
            sa = np.empty(ns + 1, dtype=np.int32)
            work = np.empty(ns, dtype=np.uint8)
            t, s_ref = reference_step(lib, text, sa, work, threads)
            whole = "the whole workload text" if ns == n else f"the first {ns} bytes of the workload text"
            cpu = {"value": ns / t / 1e6, "unit": "MB/s", "cores": threads, "cpu_model": cpu_model(), "kind": "reference",
                   "sample": f"{whole}, SA + BWT via the reference's two public calls, 1 repetition, {threads} threads; {REF_BUILD_NOTE}"}
            if ns == n and (s_ref != sentinel or not np.array_equal(work, h_work.numpy())):
                raise SystemExit("bench.py: the reference's BWT of the workload text differs from ours")
            tu = reference_unbwt(lib, work, s_ref, threads, text)
            cpu_unbwt = {"value": ns / tu / 1e6, "unit": "MB/s", "cores": threads, "kind": "reference",
                         "sample": f"reverse_burrows_wheeler_transform of the BWT of {whole} ({ns} bytes), 1 repetition, {threads} threads"}
            del text, sa, work

    # ---- the drop-in C++ path with pageable vectors (its own process; the text travels through /dev/shm)
    e2e_facade = None
    if world == 1 and not args.no_facade:
        e2e_facade = measure_facade(host_text.numpy(), args)

    # ---- inverse BWT (BASELINE.json configs[3]); frees the SA + BWT buffers first
    del d_sa, d_bwt, h_sa, h_work, d_text, host_text
    eng.release_workspace()
    torch.cuda.empty_cache()
    unbwt = None
    if world == 1 and not args.no_unbwt:
        try:
            unbwt = measure_unbwt(eng, torch, args, stream, peak)
            if unbwt is not None:
                unbwt["cpu_baseline"] = cpu_unbwt
        except Exception as exc:
            unbwt = {"error": str(exc)[:300]}

    line = {
        "metric": "sa_bwt_input_throughput", "value": value, "unit": "MB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/int32 (u64 sort keys)", "data": "synthetic",
        "config": {"workload": wl, "description": desc, "n_bytes": n, "per_gpu_bytes": n,
                   "parallelism": "1 independent text per GPU (--mode independent)" if world > 1 else "single GPU",
                   "l2": "working set (>= 44 n bytes) far larger than the 126 MB L2; no flush needed",
                   "step": "suffix array (n+1 int32) + forward BWT (n bytes + sentinel index) of the text"},
        "e2e": {"value": world * n / e2e_s / 1e6, "unit": "MB/s", "h2d_bytes_per_step": 2 * n, "d2h_bytes_per_step": 4 * (n + 1) + n + 4,
                "ms_per_step": e2e_s * 1e3, "gpu_launches_per_step": int(e2e_launches),
                "path": "b200sa_suffix_array + b200sa_bwt (the reference's two public calls; the second recognises the resident text and reuses "
                        "the sort), pinned host buffers"},
        "e2e_facade": e2e_facade,
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "kernel": "k_onesweep_pass<u64> (radix scatter sweeps)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic, "peak_source": peak_src,
                     "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": sp["alg_bytes"] / sp["launches"] if sp["launches"] else None,
                     "launches": int(sp["launches"]), "avg_launch_ms": sp["ms"] / sp["launches"] if sp["launches"] else None,
                     "share_of_step": sp["ms"] / dev_ms if dev_ms else None},
        "cpu_baseline": cpu,
        "unbwt": unbwt,
        "clocks": clocks,
        "rounds_per_step": prof["rounds"] / args.steps, "sort_passes_per_step": prof["sort_passes"] / args.steps,
        "phases": phases,
        "extras": extras,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def measure_facade(text: np.ndarray, args):
    """tools/facade_bench.cpp: the reference-shaped C++ templates on pageable std::vector storage, in their own process"""
    exe = os.path.join(ROOT, "msufsort_b200", "lib", "facade_bench")
    if not os.path.exists(exe):
        return {"error": "msufsort_b200/lib/facade_bench not built (make facade_bench)"}
    tmpdir = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else "/tmp"
    path = os.path.join(tmpdir, "b200sa_facade_bench_%d.bin" % os.getpid())
    try:
        text.tofile(path)
        out = subprocess.run([exe, path, str(max(1, min(args.steps, 5))), "2"], capture_output=True, text=True, timeout=600)
        for ln in out.stdout.splitlines():
            if ln.startswith("{"):
                return json.loads(ln)
        return {"error": (out.stdout + out.stderr)[-300:]}
    except Exception as exc:
        return {"error": str(exc)[:300]}
    finally:
        try:
            os.remove(path)
        except OSError:
            pass


def measure_unbwt(eng, torch, args, stream, peak):
    """inverse BWT of the 2^30-2 byte Markov text on one GPU: device-resident (CUDA events), end to end through
    b200sa_unbwt (pinned host buffer, in place), phases from the engine's own event spans"""
    kind, n, desc = WORKLOADS[UNBWT_WORKLOAD]
    if args.unbwt_n:
        n = args.unbwt_n
    host_text = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    np.copyto(host_text.numpy(), gen_text(kind, n))
    d_text = host_text.cuda()
    d_bwt = torch.empty(n, dtype=torch.uint8, device="cuda")
    s = eng.bwt_dev(d_text, n, d_bwt, None, stream)          # the verified forward path produces the input
    eng.release_workspace()
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(max(1, min(args.warmup, 2))):
        eng.unbwt_dev(d_bwt, n, s, d_out, stream)
    torch.cuda.synchronize()
    eng.profile_reset()
    eng.set_profiling(True)
    l0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        eng.unbwt_dev(d_bwt, n, s, d_out, stream)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    launches = (eng.launch_count() - l0) // args.steps
    prof = eng.profile()
    eng.set_profiling(False)
    if not bool(torch.equal(d_out, d_text)):
        raise SystemExit("bench.py: the inverse BWT did not restore the text")
    del d_out, d_text
    # end to end, in place on a pinned host buffer
    h_bwt = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h_bwt.copy_(d_bwt)
    h_work = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    times = []
    for it in range(1 + args.steps):
        h_work.copy_(h_bwt)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        eng.unbwt_ptr(h_work.data_ptr(), n, s)
        dt = time.perf_counter() - t0
        if it >= 1:
            times.append(dt)
    if not bool(torch.equal(h_work, host_text)):
        raise SystemExit("bench.py: the end-to-end inverse BWT did not restore the text")
    e2e_s = sum(times) / len(times)
    walk = prof["phases"]["unbwt_walk"]
    build = prof["phases"]["unbwt_build"]
    walk_ms = walk["ms"] / args.steps
    out = {
        "metric": "unbwt_input_throughput", "value": n / ms / 1e3, "unit": "MB/s", "ms_per_step": ms, "n_bytes": n, "gpu_launches": int(launches),
        "config": {"workload": UNBWT_WORKLOAD + "_unbwt", "description": "inverse BWT of the " + desc.split(" (")[0] + " (BASELINE.json configs[3]), one GPU"},
        "e2e": {"value": n / e2e_s / 1e6, "unit": "MB/s", "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": n, "d2h_bytes_per_step": n,
                "path": "b200sa_unbwt (reverse_burrows_wheeler_transform), pinned host buffer, in place"},
        "phases": {"build_ms": build["ms"] / args.steps, "walk_ms": walk_ms},
        "roofline": {"bound": "hbm", "kernel": "k_unbwt_walk + list ranking + k_unbwt_place (latency-bound: n dependent 4-byte reads)",
                     "achieved": 7 * n / (walk_ms * 1e-3) / 1e9 if walk_ms else None, "peak": peak, "unit": "GB/s",
                     "frac": 7 * n / (walk_ms * 1e-3) / 1e9 / peak if walk_ms and peak else None,
                     "algorithmic_bytes": "7 n (4 B psi read + 1 B decoded + window store / reload + 1 B final store)",
                     "sector_level_GBps": 32 * n / (walk_ms * 1e-3) / 1e9 if walk_ms else None,
                     "sector_level_frac": 32 * n / (walk_ms * 1e-3) / 1e9 / peak if walk_ms and peak else None,
                     "note": "every step of a walker is one dependent random 32-byte DRAM sector: the sector-level figure is the HBM traffic the walk causes"},
    }
    del h_bwt, h_work, d_bwt, host_text
    eng.release_workspace()
    torch.cuda.empty_cache()
    return out


def measure_extras(eng, torch, d_text, d_sa, d_bwt, n, stream):
    """LCP array and batched-blocks timings (device resident, CUDA events, best of 2 after one warm-up call)."""
    def timed(fn):
        fn()
        best = 1e30
        for _ in range(2):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best

    out = {}
    d_lcp = torch.empty(n + 1, dtype=torch.int32, device="cuda")
    ms = timed(lambda: eng.lcp_dev(d_text, n, d_sa, d_lcp, stream))
    out["lcp"] = {"ms": ms, "MBps": n / ms / 1e3, "max_lcp": int(d_lcp.max().item())}
    del d_lcp
    count = 1024
    bs = n // count
    if bs >= 1:
        offsets = np.arange(count + 1, dtype=np.int64) * bs
        total = int(offsets[-1])
        sent = [None]

        def fwd():
            sent[0] = eng.batch_dev(d_text, offsets, d_bwt, None, stream)
        ms_f = timed(fwd)
        d_back = torch.empty(total, dtype=torch.uint8, device="cuda")
        ms_i = timed(lambda: eng.unbwt_batch_dev(d_bwt, offsets, sent[0], d_back, stream))
        out["batch_1024_blocks"] = {"block_bytes": bs, "bwt_ms": ms_f, "bwt_MBps": total / ms_f / 1e3, "unbwt_ms": ms_i,
                                    "unbwt_MBps": total / ms_i / 1e3, "roundtrip_ok": bool(torch.equal(d_back, d_text[:total]))}
    return out


def run_sharded(args, wl):
    """ONE text sharded over all ranks (strong scaling): value = n / max-over-ranks time.  The line also carries configs[2]
    (2^30-2 ACGT + repeats) and configs[3] (inverse BWT of the 2^30-2 Markov text) on the same N GPUs."""
    import torch
    import torch.distributed as dist
    from msufsort_b200.api import Engine, torch_stream_handle
    from msufsort_b200.sharded import ShardedSorter

    kind, n, desc = WORKLOADS[wl]
    if args.n:
        n = args.n
    world = int(os.environ["WORLD_SIZE"]); rank = int(os.environ["RANK"]); local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this framework has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    eng = Engine(local_rank)
    sorter = ShardedSorter(eng, isa=args.isa)
    stream = torch_stream_handle()
    peak, peak_src = load_peaks()

    def barrier():
        dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        tt = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def sum_over_ranks(x: int) -> int:
        tt = torch.tensor([x], dtype=torch.int64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.SUM)
        return int(tt.item())

    def time_sharded(d_text, nn, steps, warmup, sample_clocks):
        """K sharded SA + BWT steps of d_text: (ms per step, last result, rank-0 profile, launches per step summed over ranks, clocks)"""
        for _ in range(warmup):
            res = sorter.suffix_array_bwt(d_text)
        barrier()
        eng.profile_reset(); eng.set_profiling(True)
        l0 = eng.launch_count()
        sampler = ClockSampler(local_rank)
        if sample_clocks and rank == 0:
            sampler.start()
            time.sleep(0.15)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(steps):
            res = sorter.suffix_array_bwt(d_text)
        ev1.record()
        barrier()
        ms = max_over_ranks(ev0.elapsed_time(ev1)) / steps
        clocks = sampler.stop() if sample_clocks and rank == 0 else None
        prof = eng.profile(); eng.set_profiling(False)
        launches = sum_over_ranks(eng.launch_count() - l0) // steps
        # correctness outside the timed region: assemble the SA and let the GPU validator judge it
        full_sa = sorter.gather_sa(res)
        bad = eng.check_suffix_array_dev(d_text, nn, full_sa, stream)
        if bad != 0:
            raise SystemExit(f"bench.py: sharded SA has {bad} bad rows")
        # ... and the assembled BWT must be the gather of that suffix array (sentinel row = the row of suffix 0)
        full_bwt = sorter.gather_bwt(res)
        s = res.sentinel
        if int(full_sa[s]) != 0:
            raise SystemExit("bench.py: sharded sentinel index is not the row of suffix 0")
        rows = torch.arange(0, nn + 1, device="cuda")
        rows = rows[rows != s]
        want = d_text[(full_sa[rows].long() - 1)]
        if not bool(torch.equal(full_bwt, want)):
            raise SystemExit("bench.py: sharded BWT differs from the gather of the validated suffix array")
        del full_sa, full_bwt, rows, want
        return ms, res, prof, launches, clocks

    # ---- headline: the BASELINE.json configs[1] text, sharded
    host_text = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    np.copyto(host_text.numpy(), gen_text(kind, n))            # the same text on every rank
    d_text = host_text.cuda()
    ms_per_step, res, prof, launches, clocks = time_sharded(d_text, n, args.steps, args.warmup, True)
    counts = sorter.owned_counts(res)
    nvlink_bytes = res.exchanged_bytes
    # end to end with host buffers: every rank uploads the text over its own PCIe link, the results leave as disjoint slices
    h_sa = torch.empty(n + 1, dtype=torch.int32, pin_memory=True)
    h_bwt = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    e2e_times = []
    for it in range(1 + args.steps):
        barrier()
        t0 = time.perf_counter()
        r2 = sorter.suffix_array_bwt_host(host_text, h_sa, h_bwt)
        dt = time.perf_counter() - t0
        if it >= 1:
            e2e_times.append(dt)
    e2e_s = max_over_ranks(sum(e2e_times) / len(e2e_times))
    slice_rows, slice_bytes = r2.row_end - r2.row_begin, r2.out_end - r2.out_begin
    assert bool(torch.equal(h_sa[r2.row_begin:r2.row_end].cuda(), res.sa[res.row_begin:res.row_end]))
    d2h_total = sum_over_ranks(4 * slice_rows + slice_bytes)
    del h_sa, h_bwt, d_text, host_text, res, r2
    sorter.release()
    torch.cuda.empty_cache()

    # ---- configs[2]: 2^30-2 ACGT bases with repeats on the same GPUs
    big = None
    if not args.no_big:
        try:
            k2, n2, d2 = WORKLOADS["acgt_1GiB"]
            if args.big_n:
                n2 = args.big_n
            d_big = torch.from_numpy(gen_text(k2, n2)).cuda()
            ms2, res2, prof2, launches2, _ = time_sharded(d_big, n2, args.steps, 2, False)
            big = {"workload": "acgt_1GiB", "description": d2, "n_bytes": n2, "value": n2 / (ms2 * 1e-3) / 1e6, "unit": "MB/s", "ms_per_step": ms2,
                   "rounds": res2.rounds, "owned_suffixes_per_rank": sorter.owned_counts(res2), "gpu_launches": int(launches2),
                   "nvlink_bytes_stored_by_rank0_per_step": res2.exchanged_bytes,
                   "single_gpu_ms": 120.4, "single_gpu_source": "profiles/r02_bench_acgt_1GiB_n1.json (bench.py --workload acgt_1GiB on one B200, round 2)",
                   "phases_rank0_ms": {k: v["ms"] / args.steps for k, v in prof2["phases"].items() if v["launches"]}}
            del d_big, res2
            sorter.release()
            torch.cuda.empty_cache()
        except Exception as exc:
            # a rank that fails in the middle of a collective sequence cannot rejoin its peers: the whole job stops (torchrun
            # then ends the other ranks) instead of hanging at the next barrier
            sys.stderr.write(f"bench.py: rank {rank}: sharded configs[2] run failed: {exc}\n")
            sys.stderr.flush()
            os._exit(1)

    # ---- configs[3]: inverse BWT of the 2^30-2 byte Markov text, walkers split over the ranks
    unbwt = None
    if not args.no_unbwt:
        k3, n3, d3 = WORKLOADS[UNBWT_WORKLOAD]
        if args.unbwt_n:
            n3 = args.unbwt_n
        d_t3 = torch.from_numpy(gen_text(k3, n3)).cuda()
        r3 = sorter.suffix_array_bwt(d_t3)                       # the verified forward path produces the input
        d_b3 = sorter.gather_bwt(r3)
        s3 = r3.sentinel
        del r3
        sorter.release()
        torch.cuda.empty_cache()
        peer = args.isa == "peer"
        inv = (lambda: sorter.inverse_bwt(d_b3, s3, gather_all=False)) if peer else (lambda: sorter.inverse_bwt(d_b3, s3))
        back = inv()
        barrier()
        usteps = max(1, min(args.steps, 5))
        eng.profile_reset(); eng.set_profiling(True)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(usteps):
            back = inv()
        ev1.record()
        barrier()
        ms3 = max_over_ranks(ev0.elapsed_time(ev1)) / usteps
        prof3 = eng.profile(); eng.set_profiling(False)
        # correctness outside the timed region: this rank's slice of the text, and the slices together cover the text
        if peer:
            out3, b3, e3 = back
            if not bool(torch.equal(out3[b3:e3], d_t3[b3:e3])) or sum_over_ranks(e3 - b3) != n3:
                raise SystemExit("bench.py: the sharded inverse BWT did not restore the text")
            del out3
        elif not bool(torch.equal(back, d_t3)):
            raise SystemExit("bench.py: the sharded inverse BWT did not restore the text")
        unbwt = {"metric": "unbwt_input_throughput", "value": n3 / (ms3 * 1e-3) / 1e6, "unit": "MB/s", "ms_per_step": ms3, "n_bytes": n3,
                 "config": {"workload": UNBWT_WORKLOAD + "_unbwt", "parallelism": f"psi table on every GPU, walkers split over {world} GPUs, bytes stored into the "
                            "owner of their text position over NVLink; every rank ends with its slice of the text (as it ends with its rows of the suffix array)"},
                 "phases_rank0_ms": {"build": prof3["phases"]["unbwt_build"]["ms"] / usteps, "walk": prof3["phases"]["unbwt_walk"]["ms"] / usteps}}
        del back, d_b3, d_t3

    if rank == 0:
        sp = prof["phases"]["sort_pass"]
        achieved = sp["alg_bytes"] / (sp["ms"] * 1e-3) / 1e9 if sp["ms"] > 0 else 0.0
        kernel_ms = sum(v["ms"] for v in prof["phases"].values()) / args.steps
        line = {
            "metric": "sa_bwt_input_throughput", "value": n / (ms_per_step * 1e-3) / 1e6, "unit": "MB/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u8/int32 (u64 sort keys)", "data": "synthetic",
            "config": {"workload": wl, "description": desc, "n_bytes": n,
                       "parallelism": f"one text sharded by key range over {world} GPUs, ISA {args.isa}"
                                      + (" (NVLink peer memory; round loop in C++, control plane = shared-memory barriers)" if args.isa == "peer" else " (NCCL)"),
                       "l2": "working set per GPU (>= 44 n / N bytes + the n-byte text) far larger than the 126 MB L2; no flush needed",
                       "owned_suffixes_per_rank": counts, "rounds": rounds_of(prof, args.steps)},
            "e2e": {"value": n / e2e_s / 1e6, "unit": "MB/s", "ms_per_step": e2e_s * 1e3, "h2d_bytes_per_step": n, "d2h_bytes_per_step": d2h_total,
                    "path": "ShardedSorter.suffix_array_bwt_host: every rank uploads 1/N of the pinned host text and the slices are all-gathered over NVLink, b200sa_shard_sort, every rank downloads its rows of the "
                            "suffix array and its bytes of the BWT"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_onesweep_pass<u64> (rank 0)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak if peak else None, "traffic": None, "peak_source": peak_src,
                         "launches": int(sp["launches"]), "avg_launch_ms": sp["ms"] / sp["launches"] if sp["launches"] else None},
            "nvlink": {"bytes_stored_by_rank0_per_step": nvlink_bytes,
                       "send_ms_rank0_per_step": prof["phases"]["peer_send"]["ms"] / args.steps,
                       "GBps_rank0_during_sends": nvlink_bytes / (prof["phases"]["peer_send"]["ms"] / args.steps * 1e-3) / 1e9 if prof["phases"]["peer_send"]["ms"] else None,
                       "bytes_per_round_rank0": nvlink_bytes / max(1.0, rounds_of(prof, args.steps)),
                       "peak_GBps_per_direction": 900.0,
                       "note": "new (suffix, rank) pairs stored in bulk into the owners' inboxes (k_peer_send); rank[suffix + h] is loaded from the owners' HBM by the "
                               "group-sort / key-build kernels (4-byte remote loads, not counted here)"},
            "limiter": {"kernel_ms_rank0": kernel_ms, "step_ms": ms_per_step, "kernel_share": kernel_ms / ms_per_step if ms_per_step else None,
                        "note": "step - kernels = host round trips per doubling round (counter read-backs, 2 barriers) + waiting for the slowest rank"},
            "cpu_baseline": None, "clocks": clocks,
            "phases_rank0": {k: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["launches"] / args.steps}
                             for k, v in prof["phases"].items() if v["launches"]},
            "sharded_1GiB": big,
            "unbwt": unbwt,
        }
        print(json.dumps(line))
    dist.barrier()
    sorter.close()
    eng.close()
    dist.destroy_process_group()
    return 0


def rounds_of(prof, steps):
    return prof["rounds"] / steps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="markov3_256MiB", choices=sorted(WORKLOADS))
    ap.add_argument("--n", type=int, default=0, help="override the text size (debugging)")
    ap.add_argument("--big-n", type=int, default=0, help="override the size of the sharded configs[2] text (debugging)")
    ap.add_argument("--unbwt-n", type=int, default=0, help="override the size of the inverse-BWT text (debugging)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-single-thread", action="store_true", help="reference arm: skip the numThreads = 1 repetition")
    ap.add_argument("--no-extras", action="store_true", help="skip the LCP / batched-blocks timings appended as 'extras'")
    ap.add_argument("--no-facade", action="store_true", help="skip the C++ facade end-to-end run ('e2e_facade')")
    ap.add_argument("--no-unbwt", action="store_true", help="skip the inverse-BWT record ('unbwt')")
    ap.add_argument("--no-big", action="store_true", help="N>1: skip the 2^30-2 ACGT record ('sharded_1GiB')")
    ap.add_argument("--isa", default="peer", choices=["owner", "replicated", "peer"], help="sharded mode: how the ISA travels (see msufsort_b200/sharded.py)")
    ap.add_argument("--mode", default="sharded", choices=["independent", "sharded"],
                    help="N>1 only. sharded (default): ONE text partitioned by key range over the N GPUs, the inverse suffix array in "
                         "NVLink peer memory, strong scaling (north_star item 4). independent: one text per GPU, no data-path exchange, "
                         "weak scaling (batches of independent blocks)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args, args.workload)
    if args.mode == "sharded" and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        try:
            return run_sharded(args, args.workload)
        except BaseException as exc:  # see run_sharded: never leave the peers waiting at a barrier
            if isinstance(exc, SystemExit) and exc.code in (0, None):
                raise
            import traceback
            traceback.print_exc()
            sys.stderr.flush()
            os._exit(1)
    return run_ours(args, args.workload)


if __name__ == "__main__":
    sys.exit(main())

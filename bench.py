#!/usr/bin/env python
"""bench.py -- headline benchmark of the Agilex-NTT hot path on B200.

Metric (BASELINE.json): NTT+INTT pairs/sec at n=4096, one 30-bit prime, batch of 65,536 polynomials per GPU
(configs[1]); a "step" is one forward launch + one inverse launch over the whole 1 GiB batch, in place, inputs
resident in HBM.  Multi-GPU: one process per GPU (torchrun), the batch dimension is sharded with no data-path
collective (weak scaling: 65,536 polynomials per GPU); the only collectives are the barrier and the max-over-ranks
reduction of the timing.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      # the CPU path timed on this box's host cores

Prints ONE JSON line (rank 0).  Extra keys beyond the base contract: roofline, cpu_baseline, e2e, clocks,
gpu_launches, kernels.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "NTT+INTT/sec at n=4096, 30-bit q"
UNIT = "pairs/s"
N_DEFAULT = 4096
PRIME = 1053818881                  # first 30-bit SEAL-Embedded prime
BATCH_PER_GPU = 65536               # configs[1]
SEED = 1234                         # timing seed
HBM_FALLBACK_GBS = 6650.0           # /opt/skills/guides/B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def host_threads() -> int:
    """All host threads this process may use -- not OMP_NUM_THREADS, which torchrun pins to 1."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return os.cpu_count() or 1


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


# ----------------------------------------------------------------------------------------------- clocks sampler

class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, cuda_index: int):
        self.samples, self.reasons, self.power = [], set(), []
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None
        self._h = None
        try:
            import pynvml
            import torch
            self._nv = pynvml
            pynvml.nvmlInit()
            h = None
            try:
                uuid = str(torch.cuda.get_device_properties(cuda_index).uuid)
                h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
                idx = cuda_index
                if vis and all(v.strip().isdigit() for v in vis.split(",")):
                    idx = int(vis.split(",")[cuda_index])
                h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self._h = h
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
        except Exception as e:  # NVML missing: report that instead of inventing numbers
            self.error = repr(e)

    def _loop(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self._h) / 1000.0)
            except Exception:
                pass
            self._stop.wait(0.005)

    def start(self):
        if self._h is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()

    def stop(self) -> dict:
        self._stop.set()
        if self._thr:
            self._thr.join(2.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0,
                    "note": getattr(self, "error", "no samples")}
        return {"sm_mhz": int(statistics.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples),
                "power_w_max": round(max(self.power), 1) if self.power else None}


# ------------------------------------------------------------------------------------------------ CPU baseline

def cpu_pairs_per_sec(n: int, primes, sample_polys: int, reps: int, variant: str, threads: int):
    """Time the oracle port (oracle/ntt_oracle.c) on `sample_polys` fwd+inv pairs, `reps` times; returns pairs/s."""
    from oracle import oracle as O
    P = O.Plan(n, primes)
    x = P.synthetic(sample_polys, seed=SEED)
    P.fwd(x, variant=variant, threads=threads)      # warm caches / thread pool
    P.inv(x, variant=variant, threads=threads)
    t0 = time.perf_counter()
    for _ in range(reps):
        P.fwd(x, variant=variant, threads=threads)
        P.inv(x, variant=variant, threads=threads)
    dt = time.perf_counter() - t0
    return sample_polys * len(primes) * reps / dt, dt


def reference_code_rate():
    """The reference's OWN kernel code (src/kernel/ntt.cpp compiled against the host SYCL stand-in, oracle/_ref), timed
    on the one BASELINE size it supports (n=1024, forward only, u64, one host thread -- it is a sequential emulation
    of the FPGA pipeline).  Returned as context next to the port's numbers; None where oracle/_ref was not built."""
    try:
        import numpy as np
        from oracle import oracle as O
        N, frames = 1024, 256
        if not O.ref_available(N):
            return None
        tw, pre = O.tables_u64(N, PRIME)
        x = (np.arange(N * frames, dtype=np.uint64) * np.uint64(2654435761)) % np.uint64(PRIME)
        O.reference_fwd_u64(N, x, x, PRIME, tw, pre, frames)
        t0 = time.perf_counter()
        reps = 4
        for _ in range(reps):
            O.reference_fwd_u64(N, x, x, PRIME, tw, pre, frames)
        dt = time.perf_counter() - t0
        return {"what": "reference's own fwd_ntt_kernel code via oracle/_ref, n=1024 forward u64, 1 thread",
                "forward_transforms_per_s": frames * reps / dt}
    except Exception as e:       # never let the context number break the bench line
        return {"error": repr(e)}


def run_reference(args, rank: int):
    """--impl reference: the reference's CPU implementation of the path, on this box's host cores.

    The reference's own kernel (src/kernel/ntt.cpp) is built only for N in {32,1024,8192,16384,32768} (ntt.h:11-23,
    #error otherwise), forward only, and needs the oneAPI FPGA emulator; n=4096 forward+inverse therefore runs the
    oracle port (oracle/ntt_oracle.c, Harvey/Shoup lazy butterflies = the arithmetic of ntt.cpp:331-393 at u32),
    OpenMP over polynomials on all host threads.  Each step is a bounded sample of the workload."""
    if rank != 0:
        return
    from oracle import oracle as O
    threads = host_threads()
    sample = min(args.cpu_sample, 16384)            # per step; K steps of this stay within a couple of minutes
    P = O.Plan(args.n, [PRIME])
    x = P.synthetic(sample, seed=SEED)
    for _ in range(max(1, min(args.warmup, 3))):
        P.fwd(x, threads=threads); P.inv(x, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        P.fwd(x, threads=threads)
        P.inv(x, threads=threads)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32", "data": "synthetic",
        "config": {"workload": f"n={args.n} single 30-bit prime q={PRIME}, forward+inverse NTT (configs[1]); "
                               f"each step = bounded sample of {sample} polynomials of the 65,536-polynomial batch",
                   "n": args.n, "nlimbs": 1, "batch_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} fwd+inv pairs per step x {args.steps} steps, Shoup-lazy C oracle, OpenMP",
                         "reference_code": reference_code_rate()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------- ours

def load_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic(kernel_key: str):
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json")))
        return d.get(kernel_key)
    except Exception:
        return None


def run_ours(args, rank: int, local_rank: int, world: int):
    import torch
    import torch.distributed as dist
    import agilex_ntt_b200 as A

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback (use --impl reference "
                         "for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # keep stdout to the one JSON line: NCCL prints its version banner on stdout when the communicator is
        # created, so point fd 1 at stderr while that happens
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()

    n, L, B = args.n, 1, args.batch
    if args.total_batch:                        # strong scaling (SURVEY.md s.8(d) cfg 5): a fixed global batch split over the ranks
        if args.total_batch % world:
            raise SystemExit("bench.py: --total-batch must be a multiple of the number of GPUs")
        B = args.total_batch // world
    ctx = A.Context(n, [PRIME], device=local_rank)
    data = torch.empty(B * L * n, dtype=torch.int32, device=dev)
    ctx.fill_synthetic(data, seed=SEED, first_poly=rank * B)      # shard = slice of the global synthetic batch
    chk0 = ctx.checksum(data, first_index=rank * B * L * n)
    stream = torch.cuda.current_stream()

    for _ in range(args.warmup):
        ctx.fwd(data); ctx.inv(data)
    torch.cuda.synchronize()

    K = args.steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(K)]
    sampler = ClockSampler(local_rank)
    launches0 = ctx.launch_count()
    barrier(); torch.cuda.synchronize()
    sampler.start()
    t_wall0 = time.perf_counter()
    for k in range(K):
        ev[k][0].record(stream)
        ctx.fwd(data)
        ev[k][1].record(stream)
        ctx.inv(data)
        ev[k][2].record(stream)
    torch.cuda.synchronize(); barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    launches = ctx.launch_count() - launches0

    total_ms = ev[0][0].elapsed_time(ev[K - 1][2])
    fwd_ms = [ev[k][0].elapsed_time(ev[k][1]) for k in range(K)]
    inv_ms = [ev[k][1].elapsed_time(ev[k][2]) for k in range(K)]
    ok = ctx.checksum(data, first_index=rank * B * L * n) == chk0        # K round trips leave the batch unchanged

    # ---- end-to-end through the host-pointer C ABI: pinned host buffers, H2D + kernels + D2H inside the timed region
    e2e_steps = max(1, min(K, args.e2e_steps))
    host = torch.empty(B * L * n, dtype=torch.int32).pin_memory()
    host.copy_(data)
    ctx.fwd_host(host); ctx.inv_host(host)                                # warm the pipeline (allocations)
    barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.fwd_host(host)      # returns once the spectra are back in host memory
        ctx.inv_host(host)
    torch.cuda.synchronize(); barrier()
    e2e_s = time.perf_counter() - t0
    ok_e2e = bool((host[: 64 * n].to(dev) == data[: 64 * n]).all())

    t = torch.tensor([total_ms, e2e_s * 1e3, statistics.mean(fwd_ms), statistics.mean(inv_ms),
                      0.0 if (ok and ok_e2e) else 1.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, fwd_avg, inv_avg, bad = [float(v) for v in t.tolist()]

    if rank == 0:
        peak, peak_src = load_peak()
        bytes_per_launch = 2.0 * n * 4 * B * L                      # read once + write once
        dom, dom_ms = ("ntt_fwd_kernel", fwd_avg) if fwd_avg >= inv_avg else ("ntt_inv_kernel", inv_avg)
        achieved = bytes_per_launch / (dom_ms * 1e-3) / 1e9
        variant = ctx.variant()
        # integer-multiply roofline: IMAD.HI issues at 32 lanes/clk/SM, so one butterfly holds the
        # multiply pipe for 8 of its 64 lane-clocks -> 16 butterflies/clk/SM (agx_microbench measured 14.9)
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965
        logn = n.bit_length() - 1
        bf_rate = B * L * (n // 2) * logn / (dom_ms * 1e-3) / (sms * mhz * 1e6)
        line = {
            "metric": METRIC, "value": world * B * L * K / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": K, "warmup": args.warmup, "ms_per_step": total_ms / K, "higher_is_better": True,
            "scaling": "strong" if args.total_batch else "weak", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"configs[1]: n={n} single 30-bit prime q={PRIME}, batch of {B} polynomials per GPU, "
                                   "forward launch + inverse launch in place",
                       "n": n, "nlimbs": L, "batch_per_gpu": B, "global_batch": world * B,
                       "l2": f"inputs {B * L * n * 4 >> 20} MiB per GPU > 126 MB L2 (no flush needed)",
                       "kernel_variant": variant, "parallelism": f"batch-sharded x{world}, no collective"},
            "roofline": {"bound": "hbm", "kernel": f"{dom}<{variant[6:-1]}>", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": load_traffic(dom),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": bytes_per_launch,
                         "launch_ms": dom_ms, "frac_of_8TBs": achieved / 8000.0,
                         "imad_butterflies_per_clk_per_sm": bf_rate, "imad_peak_butterflies_per_clk_per_sm": 16.0,
                         "imad_frac": bf_rate / 16.0,
                         "binding": "integer multiply pipe (IMAD.HI at half rate; 189 M transforms/s at 1965 MHz "
                                    "vs 200 M/s from measured HBM)"},
            "kernels": {"ntt_fwd_ms": fwd_avg, "ntt_inv_ms": inv_avg,
                        "fwd_transforms_per_s": B * L / (fwd_avg * 1e-3), "inv_transforms_per_s": B * L / (inv_avg * 1e-3),
                        "fwd_GBps": bytes_per_launch / (fwd_avg * 1e-3) / 1e9,
                        "inv_GBps": bytes_per_launch / (inv_avg * 1e-3) / 1e9},
            "e2e": {"value": world * B * L * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
                    "h2d_bytes_per_step": 2 * B * L * n * 4, "d2h_bytes_per_step": 2 * B * L * n * 4,
                    "steps": e2e_steps, "api": "agx_ntt_fwd_host + agx_ntt_inv_host on pinned host buffers"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "parity_in_bench": "round trip + e2e spot check " + ("ok" if bad == 0.0 else "FAILED"),
            "wall_s_timed_region": t_wall,
        }
        if world == 1 and not args.no_cpu:
            from oracle import oracle as O
            thr = host_threads()
            sample = args.cpu_sample
            reps = 8                      # 262,144 pairs ~ 12 CPU-seconds of the Shoup-lazy port
            v, dt = cpu_pairs_per_sec(n, [PRIME], sample, reps, "shoup", thr)
            vb, _ = cpu_pairs_per_sec(n, [PRIME], sample, 1, "barrett", thr)
            v1, _ = cpu_pairs_per_sec(n, [PRIME], max(256, sample // 16), 1, "shoup", 1)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": thr, "kind": "port",
                                    "sample": f"{sample} of the {B} polynomials x {reps} reps, fwd+inv, Shoup-lazy C "
                                              f"oracle (ntt.cpp:331-393 arithmetic at u32), OpenMP x{thr}; "
                                              f"{dt:.1f} s wall",
                                    "barrett_value": vb, "single_core_value": v1,
                                    "reference_code": reference_code_rate()}
        print(json.dumps(line), flush=True)
        if bad != 0.0:
            raise SystemExit("bench.py: parity check inside the bench FAILED")
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--n", "--ntt-size", dest="n", type=int, default=N_DEFAULT,
                    help="transform size (use --ntt-size under torchrun, whose own parser claims --n*)")
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU, help="polynomials per GPU")
    ap.add_argument("--total-batch", type=int, default=0,
                    help="strong scaling: fixed global batch split contiguously over the GPUs (cfg 5 uses 2^31/n)")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-sample", type=int, default=32768, help="polynomials in the bounded CPU sample")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if world == 1 and args.gpus > 1 and args.impl == "ours":
        # launched without torchrun: re-exec under torch.distributed.run, one process per GPU
        import subprocess
        port = 29500 + os.getpid() % 2000
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()

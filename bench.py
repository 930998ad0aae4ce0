#!/usr/bin/env python
"""bench.py -- headline benchmark of pfft_b200 (contract: see the task statement).

Metric (BASELINE.json): 3-D FFT GFlop/s = 5 N log2 N / t for a 1024^3 complex-to-complex
double-precision transform; one "step" = one forward (PFFT_TRANSPOSED_OUT) plus one backward
(PFFT_TRANSPOSED_IN) transform, i.e. 2 * 5 N log2 N flop.  N GPUs = N ranks on a 2-D pencil
mesh (1x1, 2x1, 2x2, 2x4), total problem size fixed ("strong" scaling).

  value        device-resident: inputs already in HBM, CUDA events around K steps, max over ranks
  e2e          same metric through the PFFT C API with pinned HOST buffers (H2D + D2H in the timed region)
  roofline     dominant kernel (one 1-D FFT pass over the local array): algorithmic bytes / CUDA-event time
  cpu_baseline numpy/pocketfft restatement (oracle, "port") on a bounded 512^3 sample, host cores
  --impl reference : the CPU arm alone (the real PFFT+FFTW-MPI cannot be built in this image:
                     no MPI, no FFTW), same metric/config keys.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MESH = {1: [1, 1], 2: [2, 1], 4: [2, 2], 8: [2, 4]}
METRIC = "3D FFT GFlop/s (5NlogN/t) c2c double, forward+backward"


def flops_per_transform(n):
    N = 1
    for x in n:
        N *= x
    return 5.0 * N * math.log2(N)


def cpu_sample_gflops(sample_n, steps, warmup):
    """Oracle port (numpy/scipy pocketfft) on the host cores: forward + backward of sample_n^3."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import scipy.fft as sf
    import pfft_oracle as po
    cores = os.cpu_count() or 1
    n = [sample_n] * 3
    x = po.synthetic_complex(n, n, [0, 0, 0]) if sample_n <= 256 else None
    if x is None:
        rng = np.random.default_rng(1234)
        x = rng.random(n) + 1j * rng.random(n)
    times = []
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        y = sf.fftn(x, workers=cores)
        z = sf.ifftn(y, workers=cores, norm="forward")
        t1 = time.perf_counter()
        if it >= warmup:
            times.append(t1 - t0)
        del y, z
    t = sum(times) / len(times)
    return 2 * flops_per_transform(n) / t / 1e9, t, cores


class ClockSampler:
    def __init__(self, index):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thr = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [v.strip() for v in out.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nm, v in zip(names, f[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thr.join(timeout=6)

    def summary(self):
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--size", type=int, default=1024, help="edge length (default: the headline 1024)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--transport", default=None)
    ap.add_argument("--mesh", default=None, help="process mesh P0xP1 (default: 1x1, 2x1, 2x2, 2x4 for 1/2/4/8 ranks)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n = [args.size] * 3
    mesh = [int(v) for v in args.mesh.split("x")] if args.mesh else MESH.get(world, [world, 1])
    config = {"workload": "%d^3 c2c fp64 forward(TRANSPOSED_OUT)+backward(TRANSPOSED_IN), mesh %dx%d" %
              (args.size, mesh[0], mesh[1]), "n": n, "mesh": mesh, "flags": "PFFT_TRANSPOSED_OUT/IN",
              "l2_policy": "arrays (>= 2 GiB per rank) far exceed the 126 MB L2; no flush needed"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        sample = 512 if args.size >= 512 else args.size
        g, t, cores = cpu_sample_gflops(sample, max(1, min(args.steps, 3)), 1)
        line = {"impl": "reference", "metric": METRIC, "value": g, "unit": "GFlop/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": g, "unit": "GFlop/s", "cores": cores, "kind": "port",
                                 "sample": "%d^3 c2c fp64 forward+backward, scipy.fft (pocketfft) workers=%d; "
                                           "PFFT+FFTW-MPI itself is not buildable here (no MPI, no FFTW)" % (sample, cores)},
                "e2e": {"value": g, "unit": "GFlop/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    import pfft_b200 as pf
    if args.transport:
        pf.set_transport(args.transport)
    pf.init()
    comm = pf.create_procmesh(mesh)
    T_OUT, T_IN = pf.TRANSPOSED_OUT, pf.TRANSPOSED_IN
    alloc, lni, lis, lno, los = pf.local_size("c2c", n, comm, T_OUT)
    cnt_in = int(np.prod(lni))
    # device-resident arrays (synthetic uniform data, generated on the device for `value`)
    gen = torch.Generator(device="cuda").manual_seed(1234 + rank)
    a = torch.rand(max(alloc, 1), 2, dtype=torch.float64, device="cuda", generator=gen) * 2 - 1
    b = torch.empty_like(a)
    fwd = pf.plan_dft(n, a, b, comm, pf.FORWARD, T_OUT | pf.DESTROY_INPUT)
    bwd = pf.plan_dft(n, b, a, comm, pf.BACKWARD, T_IN | pf.DESTROY_INPUT)
    if fwd is None or bwd is None:
        raise RuntimeError("planning failed: " + pf.last_error())

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        fwd.execute()
        bwd.execute()

    a0 = a[:cnt_in].clone()
    for _ in range(args.warmup):
        step()
        a[:cnt_in] /= float(np.prod(n))    # keep magnitudes bounded; not part of the timed region
    # round-trip sanity on the benchmark data itself
    rel = ((a[:cnt_in] - a0).norm() / a0.norm()).item()
    del a0
    launches0 = pf.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms_f, stage_ms_b = [], []
    barrier()
    with ClockSampler(local_rank) as clk:
        ev0.record()
        for _ in range(args.steps):
            fwd.execute()
            stage_ms_f.append(fwd.stage_times_ms())
            bwd.execute()
            stage_ms_b.append(bwd.stage_times_ms())
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    launches = pf.launch_count() - launches0
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = t.item()
    ms_per_step = ms_max / args.steps
    gflops = 2 * flops_per_transform(n) / (ms_per_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel: one 1-D FFT pass = read + write of the local array
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    local_elems = float(np.prod(n)) / world
    alg_pass = 2 * 16 * local_elems            # one read + one write of the local array
    desc = fwd.describe()

    def launches_of(plan, runs):
        """[(kernel name, mean ms, algorithmic bytes)] of one transform: a plane-fused pair is ONE launch
        that transforms two dimensions (4 * 16 B * local elements), its intermediate staying in L2."""
        d = plan.describe()
        mean = [sum(r[i] for r in runs) / len(runs) for i in range(len(runs[0]))]
        fi = d.get("fused_pair", -1) if d.get("fused_active", 0) else -1
        out, i = [], 0
        while i < len(mean):
            if i == fi:
                out.append(("fused_pair_kernel", mean[i] + mean[i + 1], 2 * alg_pass))
                i += 2
            else:
                out.append(("stage_%s_kernel" % d["kernels"][i], mean[i], alg_pass))
                i += 1
        return out

    per_launch = launches_of(fwd, stage_ms_f) + launches_of(bwd, stage_ms_b)
    by_kernel = {}
    for name, ms_k, b in per_launch:
        t = by_kernel.setdefault(name, [0.0, 0.0, 0])
        t[0] += ms_k
        t[1] += b
        t[2] += 1
    dom = max(by_kernel, key=lambda k: by_kernel[k][0])
    dom_ms, dom_bytes, dom_n = by_kernel[dom]
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_source": peak_src,
                "kernel": "%s, %d of the %d launches per step" % (dom, dom_n, launches // max(1, args.steps)),
                "algorithmic_bytes_per_launch": dom_bytes / dom_n, "avg_launch_ms": dom_ms / dom_n,
                "all_kernels": {k: {"launches_per_step": v[2], "ms_per_step": v[0],
                                    "achieved_gbs": v[1] / (v[0] * 1e-3) / 1e9 if v[0] > 0 else 0.0}
                                for k, v in by_kernel.items()},
                "stage_ms_forward": [sum(r[i] for r in stage_ms_f) / len(stage_ms_f) for i in range(len(stage_ms_f[0]))],
                "stage_ms_backward": [sum(r[i] for r in stage_ms_b) / len(stage_ms_b) for i in range(len(stage_ms_b[0]))]}
    # ---- HBM + NVLink roofline of the whole transform: per GPU and per transform
    # HBM bytes = 2 * d * 16 B * N / P, NVLink bytes (one direction) = sum over mesh dims of (16 B * N / P) * (P_d - 1) / P_d
    nvl_peak = 900.0          # GB/s per direction, NVLink 5 nominal; SM stores measured 707, copy engine 781 (profiles/microbench)
    hbm_bytes = 3 * alg_pass
    nvl_bytes = sum(16.0 * local_elems * (pd - 1) / pd for pd in mesh)
    t_hbm = hbm_bytes / (peak * 1e9) * 1e3
    t_nvl = nvl_bytes / (nvl_peak * 1e9) * 1e3
    t_transform = ms_per_step / 2
    combined = {"hbm_bytes_per_gpu": hbm_bytes, "nvlink_bytes_per_gpu": nvl_bytes, "hbm_peak_gbs": peak,
                "nvlink_peak_gbs": nvl_peak, "t_hbm_ms": t_hbm, "t_nvlink_ms": t_nvl,
                "t_roof_serial_ms": t_hbm + t_nvl, "t_roof_overlap_ms": max(t_hbm, t_nvl),
                "t_measured_ms": t_transform, "frac_of_serial_roofline": (t_hbm + t_nvl) / t_transform,
                "nvlink_gbs_in_exchange_stages": None}
    if world > 1 and desc["transport"] == "p2p":
        # stages in front of a real exchange push their remote chunks over NVLink inside the kernel
        rates = []
        fm = roofline_stage = [sum(r[i] for r in stage_ms_f) / len(stage_ms_f) for i in range(len(stage_ms_f[0]))]
        k = 0
        for pd in reversed(mesh):          # forward: first exchange over the last mesh dimension
            if pd > 1 and k < len(fm) and fm[k] > 0:
                rates.append(16.0 * local_elems * (pd - 1) / pd / (fm[k] * 1e-3) / 1e9)
            k += 1
        combined["nvlink_gbs_in_exchange_stages"] = rates
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch_%d" % args.size)
        except Exception:
            pass

    # ---- end to end: the same call with pinned HOST arrays (H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        bytes_local = cnt_in * 16
        ha = torch.empty(max(alloc, 1), 2, dtype=torch.float64, pin_memory=True)
        hb = torch.empty(max(alloc, 1), 2, dtype=torch.float64, pin_memory=True)
        ha.uniform_(-1, 1)
        e_steps = max(1, min(args.steps, 3))
        hp_a, hp_b = ha.data_ptr(), hb.data_ptr()
        fwd.execute(hp_a, hp_b)   # warm-up (allocates the staging buffers)
        bwd.execute(hp_b, hp_a)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            fwd.execute(hp_a, hp_b)
            bwd.execute(hp_b, hp_a)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        te = torch.tensor([(t1 - t0) / e_steps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = {"value": 2 * flops_per_transform(n) / te.item() / 1e9, "unit": "GFlop/s",
               "h2d_bytes_per_step": 2 * bytes_local, "d2h_bytes_per_step": 2 * bytes_local,
               "steps": e_steps, "note": "per rank bytes; pfft_execute_dft on pinned host arrays"}
        del ha, hb

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        g, tcpu, cores = cpu_sample_gflops(512 if args.size >= 512 else args.size, 1, 1)
        cpu = {"value": g, "unit": "GFlop/s", "cores": cores, "kind": "port",
               "sample": "512^3 c2c fp64 forward+backward, scipy.fft (pocketfft) workers=%d, 1 timed repetition" % cores}

    if rank == 0:
        line = {"metric": METRIC, "value": gflops, "unit": "GFlop/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "roundtrip_rel_err": rel, "clocks": clk.summary(), "gpu_launches": launches,
                "roofline": roofline, "roofline_hbm_nvlink": combined, "transport": desc["transport"]}
        if e2e:
            line["e2e"] = e2e
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line))
    fwd.destroy()
    bwd.destroy()
    comm.free()
    pf.finalize()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

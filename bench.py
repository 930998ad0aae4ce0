#!/usr/bin/env python
"""bench.py -- benchmark of pfft_b200 (contract: see the task statement).

Metric (BASELINE.json): 3-D FFT GFlop/s = 5 N log2 N / t for a 1024^3 complex-to-complex
double-precision transform; one "step" = one forward (PFFT_TRANSPOSED_OUT) plus one backward
(PFFT_TRANSPOSED_IN) transform, i.e. 2 * 5 N log2 N flop.  N GPUs = N ranks on a 2-D pencil
mesh (1x1, 2x1, 2x2, 2x4), total problem size fixed ("strong" scaling).

  value        device-resident: inputs already in HBM, CUDA events around K steps, max over ranks
  e2e          same metric through the C API with pinned HOST buffers (H2D + D2H in the timed region)
  roofline     dominant kernel: algorithmic bytes of its stages / CUDA-event time, vs the measured HBM peak
  spot_check   forward VALUES of the benchmark's own transform: K output coefficients against a direct
               fp64 summation of the (counter-hash generated) global input, all ranks contributing
  cpu_baseline numpy/pocketfft restatement (oracle, "port") on the host cores, bounded sample
  --impl reference : the CPU arm alone (the real PFFT+FFTW-MPI cannot be built in this image:
                     no MPI, no FFTW), same metric/config keys.
  --config K   : the other BASELINE.json configs (2: 512^3 c2c, 3: 1024^3 r2c/c2r fp32 padded in place,
                 4: 128^4 c2c on a 3-D mesh, 5: oversampled r2c 512^3 -> 768^3 + ghost cells); default 1 =
                 the headline 1024^3 c2c fp64.
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

T_IN, T_OUT, PADDED, DESTROY = 1, 2, 1 << 11, 1 << 9

# mesh per world size: 2-D pencils for 3-D data, 3-D meshes for the 4-D config
MESH2 = {1: [1, 1], 2: [2, 1], 4: [2, 2], 8: [2, 4]}
MESH3 = {1: [1, 1, 1], 2: [2, 1, 1], 4: [2, 2, 1], 8: [2, 2, 2]}
CONFIGS = {
    1: dict(name="1024^3 c2c fp64", kind="c2c", n=[1024] * 3, prec="f64", mesh=MESH2),
    2: dict(name="512^3 c2c fp64", kind="c2c", n=[512] * 3, prec="f64", mesh=MESH2),
    3: dict(name="1024^3 r2c/c2r fp32 PADDED in-place", kind="r2c", n=[1024] * 3, prec="f32", mesh=MESH2,
            padded=True, inplace=True),
    4: dict(name="128^4 c2c fp64 on a 3-D mesh", kind="c2c", n=[128] * 4, prec="f64", mesh=MESH3),
    5: dict(name="oversampled r2c/c2r fp64 512^3 -> 768^3 + ghost cells", kind="r2c", n=[768] * 3, ni=[512] * 3,
            prec="f64", mesh=MESH2, gc=dict(below=[2, 2, 0], above=[3, 3, 0])),
}
NVLINK_MEASURED, NVLINK_NOMINAL = 770.0, 900.0     # GB/s per direction (B200_PROFILING.md / NVLink 5 nominal)


def flops_per_transform(n, real=False):
    N = 1
    for x in n:
        N *= x
    return (2.5 if real else 5.0) * N * math.log2(N)


def metric_name(cfg):
    if cfg["kind"] == "c2c":
        return "3D FFT GFlop/s (5NlogN/t) c2c %s, forward+backward" % ("double" if cfg["prec"] == "f64" else "single")
    return "FFT GFlop/s (2.5NlogN/t) r2c+c2r %s, forward+backward" % ("double" if cfg["prec"] == "f64" else "single")


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port (scipy/pocketfft) on the host cores.  The reference itself needs MPI
# and FFTW-MPI, neither of which exists in the image.
# ---------------------------------------------------------------------------------------------
def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_arm(cfg, steps, warmup, budget_s):
    """forward + backward of the config on the host cores, `steps` timed steps after `warmup`.
    The edge length is reduced (and reported) only when the full size would not fit the host's memory
    or (steps + warmup) steps would not fit `budget_s`."""
    cores = host_cores()
    # torchrun exports OMP_NUM_THREADS=1 for nproc > 1 and pocketfft's pool obeys it: pin the pool size
    # explicitly, BEFORE scipy is imported, so that this arm does the same work at every --gpus value
    for v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = str(cores)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import scipy.fft as sf
    import pfft_oracle as po  # noqa: F401  (the port this arm times is the oracle's transform: scipy.fft)
    kind, full = cfg["kind"], cfg["n"]
    real = kind != "c2c"
    cdt = np.complex128 if cfg["prec"] == "f64" else np.complex64
    rdt = np.float64 if cfg["prec"] == "f64" else np.float32

    def make_input(ni):
        # uniform [-1, 1) values; the timing does not depend on them (the parity runs use the hashed input)
        rng = np.random.default_rng(1234)
        cnt = int(np.prod(ni)) * (1 if real else 2)
        v = rng.random(cnt, dtype=rdt)
        v *= 2
        v -= 1
        return v.reshape(ni) if real else v.view(cdt).reshape(ni)

    def run_once(n, x0):
        x = x0.copy()                                               # (the transforms below overwrite their input)
        ni = x0.shape
        t0 = time.perf_counter()
        if real:
            y = sf.rfftn(x, s=n, workers=cores)                     # zero-pads ni -> n at the end (pruned input)
            z = sf.irfftn(y, s=n, workers=cores, norm="forward")    # unnormalised backward
            z = z[tuple(slice(0, m) for m in ni)]
        else:
            y = sf.fftn(x, workers=cores, overwrite_x=True)
            z = sf.ifftn(y, workers=cores, norm="forward", overwrite_x=True)
        return time.perf_counter() - t0, float(abs(z.flat[0]))

    # calibration on a small cube of the same kind -> rate -> largest sample that fits the budget and the memory
    d = len(full)
    small = [64] * d if d == 4 else [128] * d
    xs = make_input(small)
    run_once(small, xs)                          # (first call creates the thread pool)
    rate = 2 * flops_per_transform(small, real) / min(run_once(small, xs)[0] for _ in range(3))
    del xs
    try:
        avail = int([l for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0].split()[1]) * 1024
    except Exception:
        avail = 32 << 30
    esize = (16 if cfg["prec"] == "f64" else 8)
    scale_choices = [1.0, 0.75, 0.5, 0.375, 0.25, 0.125]
    n, ni = full, cfg.get("ni", full)
    for sc in scale_choices:
        n = [max(8, int(round(v * sc))) for v in full]
        ni = [max(8, int(round(v * sc))) for v in cfg.get("ni", full)]
        N = 1
        for v in n:
            N *= v
        mem = 3.2 * N * esize * (0.5 if real else 1.0)
        # big transforms run slower per flop than the calibration cube (memory bound): factor 1.6 (measured 1.3)
        est = 2 * flops_per_transform(n, real) / rate * 1.6 * (steps + warmup)
        if mem < 0.8 * avail and est < budget_s:
            break
    times = []
    x0 = make_input(ni)
    for it in range(warmup + steps):
        t, _ = run_once(n, x0)
        if it >= warmup:
            times.append(t)
    t = sum(times) / len(times)
    g = 2 * flops_per_transform(n, real) / t / 1e9
    sample = "%s %s %s forward+backward, scipy.fft (pocketfft) workers=%d, %d timed steps after %d warm-up%s" % (
        "x".join(map(str, n)), "r2c/c2r" if real else "c2c", cfg["prec"], cores, steps, warmup,
        "" if n == full else " (REDUCED from %s: time/memory bound of the CPU arm)" % "x".join(map(str, full)))
    return dict(value=g, seconds_per_step=t, cores=cores, sample=sample, sample_n=n, full_size=(n == full),
                calibration_gflops=rate / 1e9)


def reference_arm(args, cfg, config):
    """`--impl reference`: rank 0 alone, every rank count does the identical single-process job."""
    if int(os.environ.get("RANK", "0")) != 0:
        return 0
    r = cpu_arm(cfg, args.steps, args.warmup, args.cpu_budget)
    line = {"impl": "reference", "metric": metric_name(cfg), "value": r["value"], "unit": "GFlop/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["seconds_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": cfg["prec"], "data": "synthetic",
            "config": dict(config, cpu_sample_n=r["sample_n"], cpu_sample_is_full_size=r["full_size"]),
            "cpu_baseline": {"value": r["value"], "unit": "GFlop/s", "cores": r["cores"], "kind": "port",
                             "calibration_gflops_small_cube": r["calibration_gflops"],
                             "sample": r["sample"] + "; PFFT+FFTW-MPI itself is not buildable here (no MPI, no FFTW), "
                                                     "this is the oracle's restatement, not the reference's performance"},
            "e2e": {"value": r["value"], "unit": "GFlop/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return 0


def nvlink_counters(index):
    """(tx, rx) bytes moved over all NVLink links of GPU `index` since boot (nvidia-smi nvlink -gt d), or None."""
    try:
        out = subprocess.run(["nvidia-smi", "nvlink", "-gt", "d", "-i", str(index)], capture_output=True, text=True, timeout=20).stdout
        tx = rx = 0
        seen = False
        if os.environ.get("PFFT_BENCH_DEBUG"):
            sys.stderr.write("nvidia-smi nvlink -gt d:\n" + out + "\n")
        for l in out.splitlines():
            l = l.strip()
            if "Data Tx" in l or "Data Rx" in l:
                kib = int(l.split(":")[-1].strip().split()[0])
                seen = True
                if "Tx" in l:
                    tx += kib * 1024
                else:
                    rx += kib * 1024
        if seen:
            return (tx, rx)
    except Exception:
        pass
    # NVML field values: data throughput counters in KiB, summed over all links (scope id = all links)
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        vals = pynvml.nvmlDeviceGetFieldValues(h, [(pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, 0xFFFFFFFF),
                                                   (pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX, 0xFFFFFFFF)])
        out = []
        for v in vals:
            if v.nvmlReturn != 0:
                return None
            out.append(int(v.value.ullVal) * 1024)
        if os.environ.get("PFFT_BENCH_DEBUG"):
            sys.stderr.write("nvml nvlink throughput counters (bytes): %r\n" % (out,))
        return tuple(out)
    except Exception as ex:
        if os.environ.get("PFFT_BENCH_DEBUG"):
            sys.stderr.write("nvml nvlink counters failed: %r\n" % (ex,))
        return None


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
            self._stop.wait(0.2)

    def __enter__(self):
        if os.environ.get("PFFT_BENCH_NO_CLOCKS"):      # (experiments: rule out the sampler as a source of jitter)
            return self
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join(timeout=6)

    def summary(self):
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ---------------------------------------------------------------------------------------------
# synthetic input: counter-based hash of the GLOBAL linear index (the same function as
# oracle/pfft_oracle.py: synthetic_complex, checked against it in tests/test_bench_helpers.py), so
# every mesh sees the same global array and single coefficients can be recomputed by direct summation
# ---------------------------------------------------------------------------------------------
def _i64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


def _lsr(z, k):
    return (z >> k) & ((1 << (64 - k)) - 1)


def splitmix64_torch(x):
    z = x + _i64(0x9E3779B97F4A7C15)
    z = (z ^ _lsr(z, 30)) * _i64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _i64(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


def synthetic_block(torch, n, local_n, local_start, real, dtype, device, seed=1234, row_pitch=None):
    """Local block (row-major, shape local_n; complex as (..., 2)) of the global synthetic array.  For real
    arrays `row_pitch` >= local_n[-1] pads the rows (padding = 0)."""
    d = len(n)
    pitch = row_pitch or local_n[-1]
    shape = list(local_n[:-1]) + [pitch] + ([] if real else [2])
    out = torch.zeros(shape, dtype=dtype, device=device)
    if min(local_n) == 0:
        return out
    inner = torch.zeros(local_n[1:], dtype=torch.int64, device=device)
    for t in range(1, d):
        shp = [1] * (d - 1)
        shp[t - 1] = local_n[t]
        inner = inner * n[t] + (torch.arange(local_n[t], device=device, dtype=torch.int64) + local_start[t]).reshape(shp)
    stride0 = 1
    for t in range(1, d):
        stride0 *= n[t]
    for i0 in range(local_n[0]):
        lin = inner + (i0 + local_start[0]) * stride0
        re = splitmix64_torch((lin * 2) ^ seed)
        v = _lsr(re, 11).to(torch.float64) * (2.0 ** -52) - 1.0
        if real:
            out[i0][..., :local_n[-1]] = v.to(dtype)
        else:
            im = splitmix64_torch((lin * 2 + 1) ^ seed)
            out[i0][..., 0] = v.to(dtype)
            out[i0][..., 1] = (_lsr(im, 11).to(torch.float64) * (2.0 ** -52) - 1.0).to(dtype)
    return out


def direct_partial_sums(torch, x, n, local_start, ks, real, sign=-1):
    """sum over the LOCAL block of x[j] * exp(sign 2 pi i k.j / n) for each of the K coefficients in `ks`
    (fp64, phases reduced exactly in integers).  x: local block, shape local_n (+ (2,) for complex)."""
    d = len(n)
    dev = x.device
    K = len(ks)
    ln = list(x.shape[:d])
    kk = torch.tensor(ks, dtype=torch.int64, device=dev)           # [K, d]

    def phases(t):
        j = torch.arange(ln[t], dtype=torch.int64, device=dev) + local_start[t]
        m = (j[:, None] * kk[None, :, t]) % n[t]                     # [l_t, K]
        ang = m.to(torch.float64) * (sign * 2.0 * math.pi / n[t])
        return torch.complex(torch.cos(ang), torch.sin(ang))

    acc = torch.zeros(K, dtype=torch.complex128, device=dev)
    W = [phases(t) for t in range(d)]
    rows = max(1, (1 << 24) // max(1, ln[-1]))                      # rows of the last dimension per chunk
    flat = x.reshape([-1, ln[-1]] + ([] if real else [2]))
    nrows = flat.shape[0]
    # weight of row r = prod_t<d-1 W_t[i_t, :]; built chunk-wise from the row's multi-index
    for r0 in range(0, nrows, rows):
        r1 = min(nrows, r0 + rows)
        blk = flat[r0:r1]
        xb = blk.to(torch.float64).to(torch.complex128) if real else torch.view_as_complex(blk.to(torch.float64).contiguous())
        part = xb @ W[d - 1]                                        # [rows, K]
        idx = torch.arange(r0, r1, device=dev, dtype=torch.int64)
        w = torch.ones(r1 - r0, K, dtype=torch.complex128, device=dev)
        for t in range(d - 2, -1, -1):
            it = idx % ln[t]
            idx = idx // ln[t]
            w = w * W[t][it]
        acc += (part * w).sum(0)
    return acc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--config", type=int, default=1, help="BASELINE.json config (1 = headline 1024^3 c2c fp64)")
    ap.add_argument("--size", type=int, default=None, help="override the edge length (experiments)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-spot", action="store_true")
    ap.add_argument("--transport", default=None)
    ap.add_argument("--mesh", default=None, help="process mesh, e.g. 2x4 (default: 1x1, 2x1, 2x2, 2x4 for 1/2/4/8 ranks)")
    ap.add_argument("--cpu-budget", type=float, default=None, help="seconds the CPU arm may take (all steps)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    cfg = dict(CONFIGS[args.config])
    if args.size:
        f = args.size / cfg["n"][0]
        cfg["n"] = [args.size] * len(cfg["n"])
        if "ni" in cfg:
            cfg["ni"] = [int(v * f) for v in cfg["ni"]]
        cfg["name"] = cfg["name"] + " (edge %d)" % args.size
    n = cfg["n"]
    ni = cfg.get("ni", n)
    d = len(n)
    real = cfg["kind"] != "c2c"
    mesh = [int(v) for v in args.mesh.split("x")] if args.mesh else cfg["mesh"].get(world, [world] + [1] * (len(cfg["mesh"][1]) - 1))
    config = {"workload": "%s forward(TRANSPOSED_OUT)+backward(TRANSPOSED_IN), mesh %s" % (cfg["name"], "x".join(map(str, mesh))),
              "n": n, "ni": ni, "mesh": mesh, "flags": "PFFT_TRANSPOSED_OUT/IN" + ("|PFFT_PADDED_R2C" if cfg.get("padded") else ""),
              "baseline_config": args.config,
              "l2_policy": "arrays far exceed the 126 MB L2; no flush needed"}
    if args.cpu_budget is None:
        args.cpu_budget = 200.0 if args.impl == "reference" else 25.0

    if args.impl == "reference":
        return reference_arm(args, cfg, config)

    import numpy as np
    import torch
    import torch.distributed as dist
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    import pfft_b200 as pf
    if args.transport:
        pf.set_transport(args.transport)
    pf.init()
    comm = pf.create_procmesh(mesh)
    rdt = torch.float64 if cfg["prec"] == "f64" else torch.float32
    npr = np.float64 if cfg["prec"] == "f64" else np.float32
    npc = np.complex128 if cfg["prec"] == "f64" else np.complex64
    esz = 8 if cfg["prec"] == "f64" else 4
    ff = T_OUT | (PADDED if cfg.get("padded") else 0)
    fb = T_IN | (PADDED if cfg.get("padded") else 0)
    if real:
        af, lni, lis, lno, los = pf.local_size("r2c", n, comm, ff, ni, n)
        ab, _, _, _, _ = pf.local_size("c2r", n, comm, fb, n, ni)
    else:
        af, lni, lis, lno, los = pf.local_size("c2c", n, comm, ff, ni, n)
        ab = af
    alloc = max(af, ab, 1)                                  # complex elements
    a = torch.zeros(alloc, 2, dtype=rdt, device=dev)
    b = a if cfg.get("inplace") else torch.zeros(alloc, 2, dtype=rdt, device=dev)
    if real:
        fwd = pf.plan_dft_r2c(n, a, b, comm, pf.FORWARD, ff | DESTROY, dtype=npr, ni=ni, no=n)
        bwd = pf.plan_dft_c2r(n, b, a, comm, pf.BACKWARD, fb | DESTROY, dtype=npr, ni=n, no=ni)
    else:
        fwd = pf.plan_dft(n, a, b, comm, pf.FORWARD, ff | DESTROY, dtype=npc, ni=ni, no=n)
        bwd = pf.plan_dft(n, b, a, comm, pf.BACKWARD, fb | DESTROY, dtype=npc, ni=n, no=ni)
    if fwd is None or bwd is None:
        raise RuntimeError("planning failed: " + pf.last_error())
    cnt_in = int(np.prod(lni))                              # elements of the input block (reals for r2c)
    log_ni = list(lni)
    if real:
        log_ni[-1] = min(lni[-1], ni[-1])                   # padded rows: logical reals per row

    def fill_input():
        blk = synthetic_block(torch, ni, log_ni, lis, real, rdt, dev, row_pitch=lni[-1] if real else None)
        flat = a.reshape(-1)
        flat[:blk.numel()] = blk.reshape(-1)
        return blk

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- forward VALUES of this very transform against a direct fp64 summation (K coefficients)
    spot = None
    x0 = fill_input()
    if not args.no_spot:
        K = 16
        rng = np.random.default_rng(99)
        ks = [[int(rng.integers(0, n[t])) for t in range(d)] for _ in range(K)]
        ks[0] = [0] * d
        ks[1] = [n[t] - 1 for t in range(d)]
        if real:
            for q in ks:
                q[-1] = q[-1] % (n[-1] // 2 + 1)
        xin = x0[..., :log_ni[-1]] if real else x0
        want = direct_partial_sums(torch, xin, n, lis, ks, real)
        if world > 1:
            wr = torch.view_as_real(want).contiguous()
            dist.all_reduce(wr)
            want = torch.view_as_complex(wr)
        del xin
    del x0
    fwd.execute()
    if not args.no_spot:
        # TRANSPOSED_OUT memory order: dims 1..r, 0, r+1..d-1 (reference doc: transposed output)
        r_pm = len(mesh) if not (d == 3 and len(mesh) == 3) else 2
        order = list(range(1, r_pm + 1)) + [0] + list(range(r_pm + 1, d))
        shp = [lno[t] for t in order]
        cnt_out = int(np.prod(shp))
        got = torch.zeros(K, dtype=torch.complex128, device=dev)
        if cnt_out > 0:
            flat_out = b.reshape(-1, 2)
            strides = [0] * d
            s = 1
            for t in reversed(order):
                strides[t] = s
                s *= lno[t]
            for q, k in enumerate(ks):
                loc = [k[t] - los[t] for t in range(d)]
                if all(0 <= loc[t] < lno[t] for t in range(d)):
                    off = sum(loc[t] * strides[t] for t in range(d))
                    e = flat_out[off].to(torch.float64)
                    got[q] = torch.complex(e[0], e[1])
        if world > 1:
            gr = torch.view_as_real(got).contiguous()
            dist.all_reduce(gr)
            got = torch.view_as_complex(gr)
        err = (got - want).abs()
        rms = float(np.sqrt(np.prod(ni)) * (1.0 / math.sqrt(3.0)) * (1.0 if real else math.sqrt(2.0)))   # rms |X_k| of uniform [-1,1) data
        spot = {"coefficients": K, "max_abs_err": float(err.max().item()), "rms_output": rms,
                "max_err_over_rms": float(err.max().item()) / rms,
                "max_rel_err": float((err / want.abs().clamp_min(1e-300)).max().item()),
                "method": "direct fp64 summation of the hashed global input vs pfft_execute output, all ranks"}

    # ---- round trip on the benchmark data itself
    bwd.execute()
    Ntot = float(np.prod(n))
    x0 = synthetic_block(torch, ni, log_ni, lis, real, rdt, dev, row_pitch=lni[-1] if real else None).reshape(-1)
    back = a.reshape(-1)[:x0.numel()].to(torch.float64) / Ntot
    if real and lni[-1] != log_ni[-1]:
        back = back.reshape(-1, lni[-1])[:, :log_ni[-1]]
        x0 = x0.reshape(-1, lni[-1])[:, :log_ni[-1]]
    num = (back - x0.to(torch.float64)).norm() ** 2
    den = x0.to(torch.float64).norm() ** 2
    if world > 1:
        nd = torch.stack([num, den])
        dist.all_reduce(nd)
        num, den = nd[0], nd[1]
    rel = math.sqrt(num.item() / max(den.item(), 1e-300))
    del back, x0

    # un-normalised pairs grow the data by prod(n) per step: pre-scale so K steps stay inside the exponent range
    growth = math.log2(Ntot)
    total_steps = args.warmup + args.steps
    fill_input()
    max_exp = 1000.0 if cfg["prec"] == "f64" else 120.0
    rescale_every = max(1, int(max_exp // growth))             # steps between re-normalisations
    a.mul_(2.0 ** (-(min(total_steps, rescale_every) * growth) / 2))

    def step(i):
        fwd.execute()
        bwd.execute()
        if (i + 1) % rescale_every == 0:
            a.mul_(2.0 ** (-rescale_every * growth))           # (one extra pass every `rescale_every` steps, counted)

    for i in range(args.warmup):
        step(i)
    launches0 = pf.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    stage_ms_f, stage_ms_b, xch_ms = [], [], []
    nvl0 = nvlink_counters(local_rank) if (world > 1 and rank == 0) else None
    with ClockSampler(local_rank) as clk:
        barrier()          # everything that could skew the ranks (sampler start, counter reads) is behind us
        ev0.record()
        for i in range(args.steps):
            fwd.execute()
            stage_ms_f.append(fwd.stage_times_ms())
            bwd.execute()
            stage_ms_b.append(bwd.stage_times_ms())
            xch_ms.append(fwd.exchange_times_ms() + bwd.exchange_times_ms())
            if (args.warmup + i + 1) % rescale_every == 0:
                a.mul_(2.0 ** (-rescale_every * growth))
        ev1.record()
        barrier()
    ms = ev0.elapsed_time(ev1)
    nvl1 = nvlink_counters(local_rank) if nvl0 else None
    launches = pf.launch_count() - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = t.item() / args.steps
    gflops = 2 * flops_per_transform(n, real) / (ms_per_step * 1e-3) / 1e9

    # ---- roofline of the dominant kernel.  Algorithmic bytes of a stage = the array it reads + the array it
    # writes, taken from the plan's own stage list (pruned / real stages count what they touch)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"])
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"

    def launches_of(plan, runs):
        dsc = plan.describe()
        mean = [sum(r[i] for r in runs) / len(runs) for i in range(len(runs[0]))]
        out = []
        for i, g in enumerate(dsc["stages"]):
            bts = g["in_elems"] * esz * (1 if g["in_real"] else 2) + g["out_elems"] * esz * (1 if g["out_real"] else 2)
            out.append((dsc["kernel_names"][i], mean[i], float(bts), g["dim"]))
        return out, dsc

    lf, desc = launches_of(fwd, stage_ms_f)
    lb, _ = launches_of(bwd, stage_ms_b)
    by_kernel = {}
    for name, ms_k, bts, _dim in lf + lb:
        e = by_kernel.setdefault(name, [0.0, 0.0, 0])
        e[0] += ms_k
        e[1] += bts
        e[2] += 1
    dom = max(by_kernel, key=lambda k: by_kernel[k][0])
    dom_ms, dom_bytes, dom_n = by_kernel[dom]
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    traffic = None
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if world == 1 and args.config == 1 and not args.size and os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get("dram_bytes_per_launch_1024")   # ncu --set full, 1 GPU, this workload
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src,
                "kernel": "%s, %d of the %d stage launches per step" % (dom, dom_n, len(lf) + len(lb)),
                "algorithmic_bytes_per_launch": dom_bytes / dom_n, "avg_launch_ms": dom_ms / dom_n,
                "all_kernels": {k: {"launches_per_step": v[2], "ms_per_step": v[0],
                                    "achieved_gbs": v[1] / (v[0] * 1e-3) / 1e9 if v[0] > 0 else 0.0}
                                for k, v in by_kernel.items()},
                "stages_forward": [{"kernel": k, "ms": m, "bytes": bt, "gbs": bt / (m * 1e-3) / 1e9 if m > 0 else 0.0, "dim": dm}
                                   for k, m, bt, dm in lf],
                "stages_backward": [{"kernel": k, "ms": m, "bytes": bt, "gbs": bt / (m * 1e-3) / 1e9 if m > 0 else 0.0, "dim": dm}
                                    for k, m, bt, dm in lb]}
    # ---- HBM + NVLink roofline of one transform, per GPU, against BOTH denominators
    hbm_bytes = sum(bt for _, _, bt, _ in lf)
    nvl_bytes = 0.0
    nvl_rates = []
    for i, g in enumerate(desc["stages"]):
        x = g["exchange"]
        if x >= 0 and desc["exchanges"][x]["nparts"] > 1:
            xc = desc["exchanges"][x]
            sent = sum(c for q, c in enumerate(xc["send_cnt"]) if q != xc["me"]) * esz * (1 if xc["elem_real"] else 2)
            nvl_bytes += sent
            if desc["transport"] == "p2p" and lf[i][1] > 0:
                nvl_rates.append(sent / (lf[i][1] * 1e-3) / 1e9)
    t_tr = ms_per_step / 2
    t_hbm = hbm_bytes / (peak * 1e9) * 1e3
    combined = {"hbm_bytes_per_gpu": hbm_bytes, "nvlink_bytes_per_gpu": nvl_bytes, "hbm_peak_gbs": peak,
                "t_hbm_ms": t_hbm, "t_measured_ms": t_tr, "nvlink_gbs_in_exchange_stages": nvl_rates or None,
                "host_ms_ordering_exchanges_per_step": [sum(r[i] for r in xch_ms) / len(xch_ms) for i in range(len(xch_ms[0]))] if xch_ms and xch_ms[0] else None}
    if nvl0 and nvl1:
        # hardware counters of rank 0's GPU around the timed loop (forward + backward per step move 2 x nvl_bytes each way)
        combined["nvlink_counters_rank0"] = {"tx_bytes_per_step": (nvl1[0] - nvl0[0]) / args.steps,
                                             "rx_bytes_per_step": (nvl1[1] - nvl0[1]) / args.steps,
                                             "algorithmic_bytes_per_step_each_way": 2 * nvl_bytes,
                                             "source": "NVLink data throughput counters (nvidia-smi nvlink -gt d / NVML field values), all links of GPU %d" % local_rank}
    for tag, pk in (("measured", NVLINK_MEASURED), ("nominal", NVLINK_NOMINAL)):
        t_nvl = nvl_bytes / (pk * 1e9) * 1e3
        combined[tag] = {"nvlink_peak_gbs": pk, "t_nvlink_ms": t_nvl, "t_roof_serial_ms": t_hbm + t_nvl,
                         "t_roof_overlap_ms": max(t_hbm, t_nvl), "frac_of_serial_roofline": (t_hbm + t_nvl) / t_tr,
                         "frac_of_overlap_roofline": max(t_hbm, t_nvl) / t_tr}

    # ---- ghost cells (config 5): exchange + reduce on the r2c input block, device-timed
    gcell = None
    if cfg.get("gc"):
        gcell = time_ghost_cells(pf, torch, cfg, comm, lni, lis, ni, rdt, dev, esz, world, dist)

    # ---- end to end: the same transforms with pinned HOST arrays (H2D + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        bytes_in = cnt_in * esz * (1 if real else 2)
        bytes_mid = int(np.prod(lno)) * esz * 2
        ha = torch.empty(alloc, 2, dtype=rdt, pin_memory=True)
        hb = ha if cfg.get("inplace") else torch.empty(alloc, 2, dtype=rdt, pin_memory=True)
        fill_input()
        ha.copy_(a)
        e_steps = max(1, min(args.steps, 3))
        hp_a, hp_b = ha.data_ptr(), hb.data_ptr()
        fwd.execute(hp_a, hp_b)   # warm-up (allocates the staging buffers)
        bwd.execute(hp_b, hp_a)
        ha.copy_(a)
        results = {}
        for mode in ("blocking", "async"):
            barrier()
            t0 = time.perf_counter()
            for _ in range(e_steps):
                if mode == "blocking":
                    fwd.execute(hp_a, hp_b)
                    bwd.execute(hp_b, hp_a)
                else:
                    fwd.execute_async(hp_a, hp_b)
                    bwd.execute_async(hp_b, hp_a)
            torch.cuda.synchronize()
            t1 = time.perf_counter()
            te = torch.tensor([(t1 - t0) / e_steps], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            results[mode] = 2 * flops_per_transform(n, real) / te.item() / 1e9
            ha.copy_(a)
        best = max(results, key=results.get)
        # PCIe ceiling of this box: pinned-memory copies alone and in both directions at once (what bounds e2e)
        pcie = None
        try:
            nb = min(1 << 30, bytes_in)
            hsrc = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
            hdst = torch.empty(nb, dtype=torch.uint8, pin_memory=True)
            dsrc = torch.empty(nb, dtype=torch.uint8, device=dev)
            ddst = torch.empty(nb, dtype=torch.uint8, device=dev)
            s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

            def rate(h2d, d2h, reps=3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(reps):
                    if h2d:
                        with torch.cuda.stream(s1):
                            ddst.copy_(hsrc, non_blocking=True)
                    if d2h:
                        with torch.cuda.stream(s2):
                            hdst.copy_(dsrc, non_blocking=True)
                torch.cuda.synchronize()
                return reps * nb / (time.perf_counter() - t0) / 1e9
            rate(True, True, 1)
            pcie = {"h2d_alone_gbs": rate(True, False), "d2h_alone_gbs": rate(False, True),
                    "both_directions_each_gbs": rate(True, True)}
            del hsrc, hdst, dsrc, ddst
        except Exception as ex:
            pcie = {"error": repr(ex)}
        e2e = {"value": results[best], "unit": "GFlop/s",
               "h2d_bytes_per_step": bytes_in + bytes_mid, "d2h_bytes_per_step": bytes_mid + bytes_in,
               "steps": e_steps, "mode": best, "blocking_gflops": results["blocking"], "async_gflops": results["async"],
               "pcie_measured": pcie,
               "pcie_bound_gflops": (2 * flops_per_transform(n, real) / ((bytes_in + bytes_mid) / (pcie["both_directions_each_gbs"] * 1e9)) / 1e9
                                     if pcie and "both_directions_each_gbs" in pcie else None),
               "note": "per rank bytes; pfft_execute_dft (blocking) / pfftb200_execute_async (stream-ordered, chunked copies "
                       "on both PCIe directions) on pinned host arrays; every step copies input and output of BOTH transforms"}
        del ha, hb

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        # the CPU arm in a process of its own (clean thread environment, memory returned afterwards)
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--config", str(args.config),
                                "--steps", "3", "--warmup", "1", "--cpu-budget", str(args.cpu_budget)] +
                               (["--size", str(args.size)] if args.size else []),
                               capture_output=True, text=True, timeout=600, env={k: v for k, v in os.environ.items()
                                                                                 if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
            cpu = json.loads(p.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as ex:      # never lose the GPU line over the CPU arm
            cpu = {"value": None, "unit": "GFlop/s", "cores": host_cores(), "kind": "port", "sample": "failed: %r" % (ex,)}

    if rank == 0:
        line = {"metric": metric_name(cfg), "value": gflops, "unit": "GFlop/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": cfg["prec"], "data": "synthetic (counter-based hash of the global index)",
                "config": config, "roundtrip_rel_l2_err": rel, "spot_check": spot, "clocks": clk.summary(),
                "gpu_launches": launches, "roofline": roofline, "roofline_hbm_nvlink": combined,
                "transport": desc["transport"], "exchange_ordering": desc.get("exchange_ordering")}
        if gcell:
            line["ghost_cells"] = gcell
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


def time_ghost_cells(pf, torch, cfg, comm, lni, lis, ni, rdt, dev, esz, world, dist):
    """pfft_exchange + pfft_reduce (reference gcell/) on the r2c input block: CUDA-event times and the halo
    bytes they are bound by (what has to cross NVLink: the halo slabs received, resp. sent back)."""
    import ctypes as C
    gb, ga = cfg["gc"]["below"], cfg["gc"]["above"]
    L = pf.lib()
    V = pf.INT * 3
    ngc, gcs = V(), V()
    L.pfft_local_size_gc_3d.restype = pf.INT
    mem = L.pfft_local_size_gc_3d(V(*lni), V(*lis), V(*gb), V(*ga), ngc, gcs)
    ngc = list(ngc)
    buf = torch.zeros(max(int(mem), 1), dtype=rdt, device=dev)
    L.pfft_plan_rgc_3d.restype = C.c_void_p
    plan = L.pfft_plan_rgc_3d(V(*ni), V(*gb), V(*ga), C.c_void_p(buf.data_ptr()), comm.handle, C.c_uint(0))
    if not plan:
        return {"error": pf.last_error()}
    own = 1
    tot = 1
    for t in range(3):
        own *= lni[t]
        tot *= ngc[t]
    reps = 5
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    L.pfft_exchange(C.c_void_p(plan))
    L.pfft_reduce(C.c_void_p(plan))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    tx = tr = 0.0
    for _ in range(reps):
        ev[0].record()
        L.pfft_exchange(C.c_void_p(plan))
        ev[1].record()
        L.pfft_reduce(C.c_void_p(plan))
        ev[2].record()
        torch.cuda.synchronize()
        tx += ev[0].elapsed_time(ev[1]) / reps
        tr += ev[1].elapsed_time(ev[2]) / reps
    L.pfft_destroy_gcplan(C.c_void_p(plan))
    halo_bytes = (tot - own) * esz
    tt = torch.tensor([tx, tr], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return {"exchange_ms": tt[0].item(), "reduce_ms": tt[1].item(), "halo_bytes_per_rank": halo_bytes,
            "block_bytes_per_rank": own * esz, "exchange_halo_gbs": halo_bytes / (tt[0].item() * 1e-3) / 1e9,
            "reduce_halo_gbs": halo_bytes / (tt[1].item() * 1e-3) / 1e9, "ngc": ngc}


if __name__ == "__main__":
    sys.exit(main())

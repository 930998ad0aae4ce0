"""tests/baseline_worker.py -- TEST INFRASTRUCTURE ONLY.

BASELINE.json configs 3-5 at their full sizes, one process per rank (`pfftrun -np 8`, all ranks
may share one GPU), checked through size-independent properties: the reference's own round trip
(pfft_init_input_* -> forward -> clear -> backward -> scale -> pfft_check_output_*,
tests/simple_check_ousam_r2c.c:68-103) and, for ghost cells, exchange against the analytic
test pattern plus conservation of the sum under reduce (the adjoint of exchange).
"""
import ctypes as C
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import pfft_b200 as pf  # noqa: E402
import pfft_oracle as po  # noqa: E402


def round_trip(cfg, comm):
    kind, n = cfg["kind"], cfg["n"]
    ni, no = cfg.get("ni", n), cfg.get("no", n)
    single = cfg.get("precision", "double") == "single"
    real = np.float32 if single else np.float64
    cplx = np.complex64 if single else np.complex128
    ff, fb = cfg["flags_forward"], cfg["flags_backward"]
    if kind == "c2c":
        af, lni, lis, lno, los = pf.local_size("c2c", n, comm, ff, ni, no)
        ab, lni2, lis2, lno2, los2 = pf.local_size("c2c", n, comm, fb, no, ni)
        alloc = max(af, ab, 1)
        a = pf.ManagedArray(alloc, cplx)
        b = pf.ManagedArray(alloc, cplx)
        fwd = pf.plan_dft(n, a, b, comm, pf.FORWARD, ff | pf.DESTROY_INPUT, dtype=cplx, ni=ni, no=no)
        bwd = pf.plan_dft(n, b, a, comm, pf.BACKWARD, fb | pf.DESTROY_INPUT, dtype=cplx, ni=no, no=ni)
        pattern = "complex"
    else:
        af, lni, lis, lno, los = pf.local_size("r2c", n, comm, ff, ni, n)
        ab, _, _, lno2, los2 = pf.local_size("c2r", n, comm, fb, n, ni)
        alloc = max(af, ab, 1)
        a = pf.ManagedArray(2 * alloc, real)
        b = a if cfg.get("inplace") else pf.ManagedArray(alloc, cplx)
        fwd = pf.plan_dft_r2c(n, a, b, comm, pf.FORWARD, ff | pf.DESTROY_INPUT, dtype=real, ni=ni, no=n)
        bwd = pf.plan_dft_c2r(n, b, a, comm, pf.BACKWARD, fb | pf.DESTROY_INPUT, dtype=real, ni=n, no=ni)
        pattern = "real"
    if fwd is None or bwd is None:
        return dict(error="planning failed: " + pf.last_error())
    pf.init_input(pattern, ni, lni, lis, a, dtype=real)
    out = dict(error="", local_ni=lni, local_i_start=lis, local_no=lno, kernels_forward=fwd.describe()["kernels"])
    if cfg.get("gc") and pattern == "real":
        out["gc"] = ghost_cells(cfg, comm, a, ni, lni, lis)
        pf.init_input(pattern, ni, lni, lis, a, dtype=real)
    fwd.execute()
    if b is not a:      # (in place the array holds the forward result)
        pf.clear_input(pattern, ni, lni, lis, a, dtype=real)
    bwd.execute()
    out["stage_ms_forward"], out["stage_ms_backward"] = fwd.stage_times_ms(), bwd.stage_times_ms()
    out["kernels_backward"] = bwd.describe()["kernels"]
    cnt = int(np.prod(lni))
    a.array[:cnt] /= float(np.prod(n))
    out["maxerror"] = pf.check_output(pattern, ni, lni, lis, a, comm, dtype=real)
    if cfg.get("time_pairs"):
        # timing (profiles/time_baseline_configs.py): one untimed pair brings the managed arrays back to the
        # GPU, then `time_pairs` forward+backward pairs, wall clock between barriers, max over ranks
        import time
        fwd.execute()
        bwd.execute()
        comm.barrier()
        t0 = time.perf_counter()
        for _ in range(cfg["time_pairs"]):
            fwd.execute()
            bwd.execute()
        comm.barrier()
        out["ms_per_pair"] = comm.allreduce_max((time.perf_counter() - t0) / cfg["time_pairs"] * 1e3)
        out["stage_ms_forward"], out["stage_ms_backward"] = fwd.stage_times_ms(), bwd.stage_times_ms()
    fwd.destroy()
    bwd.destroy()
    a.free()
    if b is not a:
        b.free()
    return out


def ghost_cells(cfg, comm, a, n, lni, lis):
    """pfft_plan_rgc_3d on the r2c input block (already holding the test pattern): exchange, compare the
    whole ghost-cell block with the analytic pattern at the wrapped global indices; reduce, compare sums."""
    gb, ga = cfg["gc"]["below"], cfg["gc"]["above"]
    L = pf.lib()
    V = pf.INT * 3
    ngc, gcs = V(), V()
    L.pfft_local_size_gc_3d.restype = pf.INT
    mem = L.pfft_local_size_gc_3d(V(*lni), V(*lis), V(*gb), V(*ga), ngc, gcs)
    ngc, gcs = list(ngc), list(gcs)
    assert mem <= a.array.size, (mem, a.array.size)
    L.pfft_plan_rgc_3d.restype = C.c_void_p
    plan = L.pfft_plan_rgc_3d(V(*n), V(*gb), V(*ga), C.c_void_p(a.ptr), comm.handle, C.c_uint(0))
    if not plan:
        return dict(error="gc plan is NULL: " + pf.last_error())
    own_sum = float(np.sum(a.array[:int(np.prod(lni))], dtype=np.float64))
    L.pfft_exchange(C.c_void_p(plan))
    tot = int(np.prod(ngc))
    ex = np.array(a.array[:tot], copy=True).reshape(ngc)
    # local_ni may be padded in the last dimension (none here): pattern at wrapped global indices, real part
    grids = np.meshgrid(*[(np.arange(ngc[t], dtype=np.int64) + gcs[t]) % n[t] for t in range(3)], indexing="ij")
    want = po._init_scalar(n, grids).real
    ex_err = float(np.abs(ex - want).max())
    ex_sum = float(np.sum(ex, dtype=np.float64))
    L.pfft_reduce(C.c_void_p(plan))
    cnt = int(np.prod(lni))
    red_sum = float(np.sum(a.array[:cnt], dtype=np.float64))
    tail_zero = bool(np.all(a.array[cnt:tot] == 0))
    L.pfft_destroy_gcplan(C.c_void_p(plan))
    return dict(error="", ngc=ngc, gc_start=gcs, mem=int(mem), exchange_maxerr=ex_err, exchanged_sum=ex_sum,
                reduced_sum=red_sum, own_sum=own_sum, tail_zero=tail_zero)


def main():
    cfg = json.load(open(sys.argv[1]))
    outdir = sys.argv[2]
    pf.init()
    comm = pf.create_procmesh(cfg["np"])
    res = round_trip(cfg, comm)
    json.dump(res, open(os.path.join(outdir, "rank%d.json" % comm.rank), "w"))
    comm.free()
    pf.finalize()


if __name__ == "__main__":
    main()

"""tests/gpu_worker.py -- TEST INFRASTRUCTURE ONLY.

Runs one parity case through the C-ABI library on the GPU: scatter the global input,
plan, execute, and hand back every rank's output block.  Used in-process for 1-rank
cases and as `pfftrun -np N python tests/gpu_worker.py case.json outdir` for N ranks
(all ranks may share one GPU: the p2p transport maps peers through CUDA IPC).
No torch import here: buffers come from pfft_alloc_* (managed) or numpy (host).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import cases  # noqa: E402
import pfft_b200 as pf  # noqa: E402


def _dtypes(case):
    single = case.get("precision", "double") == "single"
    real = np.float32 if single else np.float64
    cplx = np.complex64 if single else np.complex128
    kind = case["kind"]
    din = real if kind in ("r2c", "r2r") else cplx
    dout = real if kind in ("c2r", "r2r") else cplx
    return real, din, dout


def run_case(case, comm, seed=0):
    """Returns dict(local_ni=..., local_no=..., local_o_start=..., out=flat ndarray, ...)."""
    kind, n, flags = case["kind"], case["n"], case.get("flags", 0)
    hm = case.get("howmany", 1)
    real, din, dout = _dtypes(case)
    r = len(case["np"])
    alloc, lni, lis, lno, los = pf.local_size(kind, n, comm, flags, case.get("ni"), case.get("no"), hm,
                                              case.get("iblock"), case.get("oblock"))
    xg = cases.make_global_input(case, seed)
    mine = cases.local_input(case, xg, lni, lis, r)
    mode = case.get("memory", "managed")
    nin_elems = max(mine.size, 1)
    nout_elems = max(int(np.prod(lno)) * hm, 1)
    inplace = case.get("inplace", False)
    if mode == "managed":
        cnt_c = alloc
        # alloc_local counts complex elements for c2c/r2c/c2r, reals for r2r
        bytes_needed = cnt_c * (np.dtype(real).itemsize * (1 if kind == "r2r" else 2))
        ma_in = pf.ManagedArray(max(bytes_needed // np.dtype(din).itemsize, nin_elems), din)
        ma_out = ma_in if inplace else pf.ManagedArray(max(bytes_needed // np.dtype(dout).itemsize, nout_elems), dout)
        a_in, a_out = ma_in.array, (ma_out.array if not inplace else np.frombuffer(ma_in.array, dtype=np.uint8).view(dout))
        h_in, h_out = ma_in, ma_out
    else:   # plain host memory: staged through HBM inside execute
        a_in = np.zeros(nin_elems, dtype=din)
        a_out = np.zeros(nout_elems, dtype=dout)
        h_in, h_out = a_in, a_out
    a_in[:mine.size] = np.nan_to_num(mine, nan=777.0).astype(din)
    planner = {"c2c": pf.plan_dft, "r2c": pf.plan_dft_r2c, "c2r": pf.plan_dft_c2r, "r2r": pf.plan_r2r}[kind]
    kw = dict(ni=case.get("ni"), no=case.get("no"), howmany=hm, iblock=case.get("iblock"), oblock=case.get("oblock"),
              skip=case.get("skip"))
    if kind == "r2r":
        plan = planner(n, h_in, h_out, comm, case["kinds"], flags, dtype=real, **kw)
    elif kind == "c2c":
        plan = planner(n, h_in, h_out, comm, case.get("sign", -1), flags, dtype=din, **kw)
    else:
        plan = planner(n, h_in, h_out, comm, case.get("sign", -1 if kind == "r2c" else +1), flags, dtype=real, **kw)
    if plan is None:
        return dict(error=pf.last_error())
    saved_in = np.array(a_in[:mine.size], copy=True)
    for _ in range(case.get("repeat", 1)):
        plan.execute()
    out = np.array(a_out[:nout_elems], copy=True)
    preserved = bool(np.array_equal(np.asarray(a_in[:mine.size]).view(np.uint8), saved_in.view(np.uint8))) \
        if not inplace else True
    desc = plan.describe()
    plan.destroy()
    if mode == "managed":
        ma_in.free()
        if not inplace:
            ma_out.free()
    dev = pf.lib().pfftb200_get_device()
    return dict(error="", local_ni=lni, local_i_start=lis, local_no=lno, local_o_start=los, out=out, device=int(dev),
                input_preserved=preserved, kernels=desc["kernels"], kernel_names=desc.get("kernel_names", []),
                transport=desc["transport"],
                fused=desc.get("fused_active", 0))


def check_case(case, results, l2=False):
    """results: list over ranks of run_case dicts -> relative max error vs the oracle; with l2=True
    (relative max error, relative L2 error of the whole gathered array -- the north_star's metric)."""
    xg = cases.make_global_input(case, 0)
    want = cases.oracle_output(case, xg)
    scale = max(1e-300, float(np.abs(want).max()))
    r = len(case["np"])
    err, ssq = 0.0, 0.0
    for res in results:
        m, s2 = cases.compare_local_output(case, want, res["out"], res["local_no"], res["local_o_start"], r, sumsq=True)
        err = max(err, m)
        ssq += s2
    if l2:
        return err / scale, float(np.sqrt(ssq) / max(1e-300, np.linalg.norm(np.asarray(want, dtype=np.complex128).ravel())))
    return err / scale


def run_gc_case(case, comm):
    """Ghost cells: exchange, snapshot, reduce.  Returns both snapshots (flat)."""
    import ctypes as C
    n, gb, ga = case["n"], case["gc_below"], case["gc_above"]
    cplx = case.get("complex", True)
    dtype = np.complex128 if cplx else np.float64
    _, lni, lis, _, _ = pf.local_size("c2c", n, comm, 0)
    L = pf.lib()
    V = pf.INT * 3
    ngc, gcs = V(), V()
    L.pfft_local_size_gc_3d.restype = pf.INT
    mem = L.pfft_local_size_gc_3d(V(*lni), V(*lis), V(*gb), V(*ga), ngc, gcs)
    ngc, gcs = list(ngc), list(gcs)
    buf = pf.ManagedArray(max(mem, int(np.prod(lni)), 1), dtype)
    rng = np.random.default_rng(5)
    xg = rng.standard_normal(n) + (1j * rng.standard_normal(n) if cplx else 0)
    mine = xg[tuple(slice(lis[t], lis[t] + lni[t]) for t in range(3))].astype(dtype)
    buf.array[:mine.size] = mine.reshape(-1)
    fn = L.pfft_plan_cgc_3d if cplx else L.pfft_plan_rgc_3d
    fn.restype = C.c_void_p
    plan = fn(V(*n), V(*gb), V(*ga), C.c_void_p(buf.ptr), comm.handle, C.c_uint(case.get("gc_flags", 0)))
    if not plan:
        return dict(error="plan is NULL: " + pf.last_error())
    L.pfft_exchange(C.c_void_p(plan))
    exchanged = np.array(buf.array[:int(np.prod(ngc))], copy=True)
    L.pfft_reduce(C.c_void_p(plan))
    reduced = np.array(buf.array[:int(np.prod(ngc))], copy=True)
    L.pfft_destroy_gcplan(C.c_void_p(plan))
    buf.free()
    return dict(error="", local_n=lni, local_start=lis, ngc=ngc, gc_start=gcs, out=np.concatenate([exchanged, reduced]))


def main():
    case = json.load(open(sys.argv[1]))
    outdir = sys.argv[2]
    pf.init()
    comm = pf.create_procmesh(case["np"])
    res = run_gc_case(case, comm) if case["kind"] == "gc" else run_case(case, comm)
    rank = comm.rank
    if res["error"]:
        json.dump(dict(error=res["error"]), open(os.path.join(outdir, "rank%d.json" % rank), "w"))
    else:
        np.save(os.path.join(outdir, "rank%d.npy" % rank), res["out"])
        meta = {k: v for k, v in res.items() if k != "out"}
        json.dump(meta, open(os.path.join(outdir, "rank%d.json" % rank), "w"))
    comm.free()
    pf.finalize()


if __name__ == "__main__":
    main()

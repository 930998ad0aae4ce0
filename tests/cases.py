"""tests/cases.py -- TEST INFRASTRUCTURE ONLY: shared helpers that scatter a global
input array over a (real or virtual) mesh, gather the outputs and compare with the
oracle (oracle/pfft_oracle.py).  Used by the CPU schedule-simulation tests and by the
GPU parity tests, so both check the same thing."""
import numpy as np

import pfft_oracle as po

KIND = {"c2c": po.C2C, "r2c": po.R2C, "c2r": po.C2R, "r2r": po.R2R}


def make_global_input(case, seed=0):
    """Global input of a case (dict with kind, n, [ni], [howmany]) as the oracle wants it."""
    kind, n = case["kind"], case["n"]
    ni = case.get("ni") or n
    hm = case.get("howmany", 1)
    d = len(n)
    rng = np.random.default_rng(seed)
    shape = list(ni) + ([hm] if hm > 1 else [])
    if kind == "c2c":
        return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    if kind in ("r2c", "r2r"):
        return rng.standard_normal(shape)
    xr = rng.standard_normal(shape)          # c2r: a Hermitian-consistent half spectrum
    return np.fft.fftn(np.fft.rfft(xr, axis=d - 1), axes=list(range(d - 1)))


def local_input(case, xg, lni, lis, rnk_pm):
    """The rank's input block, flattened in the user's memory order (padding = NaN)."""
    kind, n, flags = case["kind"], case["n"], case.get("flags", 0)
    ni = case.get("ni") or n
    d = len(n)
    si = bool(flags & po.SHIFTED_IN)
    shift = [ni[t] // 2 if si else 0 for t in range(d)]
    ln = list(lni)
    whole_rows = rnk_pm < d      # (3-D data on a 3-D mesh distributes the rows of reals as well)
    if kind == "r2c" and whole_rows:
        ln[-1] = ni[-1]
    blk = po.extract_block(xg, ln, lis, None, shift)
    if kind == "r2c" and whole_rows and lni[-1] != ni[-1]:
        pad = [(0, 0)] * blk.ndim
        pad[d - 1] = (0, lni[-1] - ni[-1])
        blk = np.pad(blk, pad, constant_values=np.nan)
    if flags & po.TRANSPOSED_IN:
        order = po.mem_order(d, rnk_pm, True)
        blk = np.ascontiguousarray(np.transpose(blk, order + list(range(d, blk.ndim))))
    return blk.reshape(-1)


def oracle_output(case, xg):
    kind, n = case["kind"], case["n"]
    d = len(n)
    skip = case.get("skip")
    r = len(case["np"])
    skip_dims = None if skip is None else [skip[min(t, r)] for t in range(d)]
    return po.global_transform(KIND[kind], xg, n, case.get("ni"), case.get("no"), case.get("sign", -1),
                               case.get("flags", 0), case.get("kinds"), skip_dims)


def compare_local_output(case, want, flat, lno, los, rnk_pm, sumsq=False):
    """max abs error of one rank's flattened output block against the oracle's global result
    (with sumsq=True: (max abs error, sum of squared errors) -- for the relative L2 error of the whole array)."""
    kind, n, flags = case["kind"], case["n"], case.get("flags", 0)
    no = case.get("no") or n
    hm = case.get("howmany", 1)
    d = len(n)
    so = bool(flags & po.SHIFTED_OUT)
    tout = bool(flags & po.TRANSPOSED_OUT)
    shift = [no[t] // 2 if so else 0 for t in range(d)]
    order = po.mem_order(d, rnk_pm, True) if tout else list(range(d))
    shp = [lno[t] for t in order] + ([hm] if hm > 1 else [])
    if int(np.prod(shp)) == 0:
        return (0.0, 0.0) if sumsq else 0.0
    got = np.asarray(flat)[:int(np.prod(shp))].reshape(shp)
    ln = list(lno)
    if kind == "c2r" and lno[-1] != no[-1] and rnk_pm < d:      # padded real rows: compare the logical part
        assert order[-1] == d - 1 or hm > 1
        got = got[..., :no[-1]] if hm == 1 else got[..., :no[-1], :]
        ln[-1] = no[-1]
    inv = np.argsort(order).tolist()
    got = np.transpose(got, inv + list(range(d, got.ndim)))
    ref = po.extract_block(want, ln, los, None, shift)
    if not ref.size:
        return (0.0, 0.0) if sumsq else 0.0
    diff = np.abs(got - ref).astype(np.float64)
    return (float(diff.max()), float((diff ** 2).sum())) if sumsq else float(diff.max())

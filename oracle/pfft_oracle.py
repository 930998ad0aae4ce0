"""oracle/pfft_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement (numpy, fp64) of what the reference computes on the hot path
`pfft_local_size_* / pfft_plan_* / pfft_execute`.  Only
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product (pfft_b200/) never does.

Pinning status
--------------
* Integer layer (block decomposition, local_ni/local_i_start/local_no/
  local_o_start, ghost-cell sizes, init/check pattern): PINNED -- checked
  against golden vectors captured from the reference's own compiled code
  (oracle/_ref/libpfft_refint.so, tests/golden/*.json, generator
  tests/golden/gen_golden.py) and, when the prebuilt file is present, against it
  live.
* Transform values: the reference has no FFT arithmetic of its own; every 1-D
  transform is an FFTW plan (third-party dependency, NOT under /root/reference;
  pinned >= 3.3.3 by pfft.pc.in:10, CI uses 3.3.4, conf/travis-install-fftw.sh:4).
  FFTW's published definition is the unnormalised DFT
  Y[k] = sum_j X[j] exp(sign * 2 pi i j k / n)  (doc/features.tex:140-156,
  doc/develop.tex:33-36), restated here with numpy's pocketfft.  The reference's
  own tests pin only the forward->backward round trip (tests/simple_check_*.c,
  tol 1e-12, tests/run_checks.sh:75) and c2r-vs-c2c consistency; forward values
  are therefore "parity unpinned" by the reference and anchored on the DFT
  definition (a brute-force O(n^2) DFT cross-check lives in tests/).

Every function cites the reference file:line it follows.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

# ---- public flag values (api/pfft.h:528-572) --------------------------------
TRANSPOSED_NONE = 0
TRANSPOSED_IN = 1 << 0
TRANSPOSED_OUT = 1 << 1
SHIFTED_NONE = 0
SHIFTED_IN = 1 << 2
SHIFTED_OUT = 1 << 3
ESTIMATE = 1 << 4
PRESERVE_INPUT = 1 << 8
DESTROY_INPUT = 1 << 9
PADDED_R2C = 1 << 11
PADDED_C2R = PADDED_R2C
FORWARD, BACKWARD = -1, +1
DEFAULT_BLOCK = 0
GC_TRANSPOSED_NONE, GC_TRANSPOSED = 0, 1

# r2r kinds = FFTW's enum values (api/pfft.h:45-55)
R2HC, HC2R, DHT, REDFT00, REDFT01, REDFT10, REDFT11, RODFT00, RODFT01, RODFT10, RODFT11 = range(11)

C2C, R2C, C2R, R2R = "c2c", "r2c", "c2r", "r2r"


# ---- a1: 1-D block decomposition (kernel/block.c:54-95) -----------------------
def num_blocks(n, blk):
    return (n + blk - 1) // blk


def global_block_size(n, user_blk, nprocs):
    """kernel/block.c:75-95: user block or ceil(n / P)."""
    return (n + nprocs - 1) // nprocs if user_blk == DEFAULT_BLOCK else user_blk


def local_block_size(n, blk, which):
    nb = num_blocks(n, blk)
    if which >= nb:
        return 0
    return n - which * blk if which == nb - 1 else blk


def local_block_offset(n, blk, which):
    """offset is 0 (not n) for ranks past the last block, kernel/block.c:54-61."""
    return 0 if which >= num_blocks(n, blk) else which * blk


def physical_size(n, kind):
    """util/util.c:207-227: last dim n/2+1 for r2c / c2r."""
    pn = list(n)
    if kind in (R2C, C2R):
        pn[-1] = n[-1] // 2 + 1
    return pn


# ---- mesh helpers (kernel/procmesh.c) -----------------------------------------
def cart_coords(np_, pid):
    """row-major rank -> coords, last mesh dim fastest (MPI_Cart_coords)."""
    c = [0] * len(np_)
    for t in range(len(np_) - 1, -1, -1):
        c[t] = pid % np_[t]
        pid //= np_[t]
    return c


def cart_rank(np_, coords):
    r = 0
    for t in range(len(np_)):
        r = r * np_[t] + coords[t]
    return r


def factorize_equal(p0, p1, q):
    """kernel/procmesh.c:367-391 incl. its initial-value quirk (opt = (1, q)
    paired with the error of (q, 1))."""
    opt_q0, opt_q1 = 1, q
    min_err = abs(p0 * q - p1 * 1.0)
    q1 = 1
    while q1 <= math.sqrt(q):
        q0 = q // q1
        if q0 * q1 == q:
            err = abs(p0 * q0 - p1 * q1)
            if err < min_err:
                min_err, opt_q0, opt_q1 = err, q0, q1
        q1 += 1
    return opt_q0, opt_q1


def coords_3dto2d(q0, q1, c3):
    """kernel/procmesh.c:191-198."""
    return [c3[0] * q0 + c3[2] // q1, c3[1] * q1 + c3[2] % q1]


def default_block_size_3dto2d(n, p0, p1, q0, q1):
    """kernel/remap_3dto2d.c:437-457."""
    oblk = [global_block_size(n[0], 0, p0 * q0), global_block_size(n[1], 0, p1 * q1), n[2]]
    iblk = [oblk[0] * q0, oblk[1] * q1, global_block_size(n[2], 0, q0 * q1)]
    mblk = [oblk[0] * q0, oblk[1], iblk[2] * q1]
    return iblk, mblk, oblk


# ---- a2/a3: local blocks of a parallel transform -------------------------------
def _evaluate_blocks(kind, rnk_n, ni, no, iblock, oblock, np_pm, transp):
    """kernel/partrafo.c:652-701.  Blocks come from PHYSICAL sizes; the middle
    (transposed) layout uses dims shifted by one (pnm+1)."""
    r = len(np_pm)
    pni, pno = physical_size(ni, kind), physical_size(no, kind)
    pnm = pni if (kind == C2R or transp & TRANSPOSED_IN) else pno
    mblock = None
    if transp & TRANSPOSED_IN:
        mblock = iblock
    if transp & TRANSPOSED_OUT:
        mblock = oblock
    ev = lambda pn, ub: [global_block_size(pn[t], DEFAULT_BLOCK if ub is None else ub[t], np_pm[t])
                         for t in range(r)]
    iblk = mblk = oblk = [0] * r
    if not transp & TRANSPOSED_IN:
        iblk = ev(pni, iblock)
    mblk = ev(pnm[1:], mblock)
    if not transp & TRANSPOSED_OUT:
        oblk = ev(pno, oblock)
    return iblk, mblk, oblk


def _decompose(n, kind, blk, coords, transposed):
    """kernel/partrafo-transposed.c:363-411."""
    pn = physical_size(n, kind)
    ln, ls = list(pn), [0] * len(pn)
    off = 1 if transposed else 0
    for t in range(len(coords)):
        ln[t + off] = local_block_size(pn[t + off], blk[t], coords[t])
        ls[t + off] = local_block_offset(pn[t + off], blk[t], coords[t])
    return ln, ls


def _fix_real_count(kind, padded, n, ln, which):
    """kernel/partrafo-transposed.c:67-85: the user interface counts REALS in
    the last dim of the real side of r2c (input) / c2r (output)."""
    if (kind == R2C and which == "in") or (kind == C2R and which == "out"):
        ln[-1] = ln[-1] * 2 if padded else n[-1]


def local_block(kind, ni, no, np_, pid, flags=0, iblock=None, oblock=None):
    """Restates PX(local_block_partrafo), kernel/partrafo.c:99-199.
    Returns (local_ni, local_i_start, local_no, local_o_start)."""
    rnk_n = len(ni)
    transp = flags & (TRANSPOSED_IN | TRANSPOSED_OUT)
    padded = bool(flags & PADDED_R2C)
    remap3d = rnk_n == 3 and len(np_) == 3
    if remap3d:
        q0, q1 = factorize_equal(np_[0], np_[1], np_[2])
        c3 = cart_coords(np_, pid)
        coords = coords_3dto2d(q0, q1, c3)
        np_pm = [np_[0] * q0, np_[1] * q1]
    else:
        coords = cart_coords(np_, pid)
        np_pm = list(np_)
    iblk, mblk, oblk = _evaluate_blocks(kind, rnk_n, ni, no, iblock, oblock, np_pm, transp)

    # init_param_size_and_trafo_flags (kernel/partrafo.c:735-834), called with n=ni
    ni_to, no_to, kind_to = list(ni), list(no), kind
    ni_ti, no_ti, kind_ti = list(no), list(no), kind
    if transp & TRANSPOSED_IN:
        ni_ti, no_ti = list(ni), list(no)
    if kind == R2C:
        # backward half runs c2c on the physical (complex) size
        ni_ti = no_ti = list(no[:-1]) + [no[-1] // 2 + 1]
        kind_ti = C2C
    if kind == C2R:
        ni_to = no_to = list(ni[:-1]) + [ni[-1] // 2 + 1]
        kind_to = C2C
        ni_ti, no_ti = list(ni), list(no)

    lni = lis = lno = los = None
    if not transp & TRANSPOSED_IN:
        a_n, a_s = _decompose(ni_to, kind_to, iblk, coords, False)
        b_n, b_s = _decompose(no_to, kind_to, mblk, coords, True)
        _fix_real_count(kind_to, padded, ni_to, a_n, "in")
        _fix_real_count(kind_to, padded, no_to, b_n, "out")
        lni, lis = a_n, a_s
        if transp & TRANSPOSED_OUT:
            lno, los = b_n, b_s
        if remap3d:
            pn = ni_to  # logical sizes: r2c input is handled like r2r (remap_3dto2d.c:92-94)
            ib3, _, ob3 = default_block_size_3dto2d(pn, np_[0], np_[1], q0, q1)
            lni = [local_block_size(pn[t], ib3[t], c3[t]) for t in range(3)]
            lis = [local_block_offset(pn[t], ib3[t], c3[t]) for t in range(3)]
    if not transp & TRANSPOSED_OUT:
        a_n, a_s = _decompose(ni_ti, kind_ti, mblk, coords, True)
        b_n, b_s = _decompose(no_ti, kind_ti, oblk, coords, False)
        _fix_real_count(kind_ti, padded, ni_ti, a_n, "in")
        _fix_real_count(kind_ti, padded, no_ti, b_n, "out")
        lno, los = b_n, b_s
        if transp & TRANSPOSED_IN:
            lni, lis = a_n, a_s
        if remap3d:
            pn = no_ti
            ib3, _, ob3 = default_block_size_3dto2d(pn, np_[0], np_[1], q0, q1)
            lno = [local_block_size(pn[t], ib3[t], c3[t]) for t in range(3)]
            los = [local_block_offset(pn[t], ib3[t], c3[t]) for t in range(3)]
    # SHIFTED start offsets (kernel/partrafo.c:178-190)
    if flags & SHIFTED_IN:
        lis = [lis[t] - ni[t] // 2 for t in range(rnk_n)]
    if flags & SHIFTED_OUT:
        los = [los[t] - no[t] // 2 for t in range(rnk_n)]
    return lni, lis, lno, los


def local_size_gc(local_n, local_start, howmany, gc_below, gc_above):
    """gcell/gcells_plan.c:51-76: (mem, local_ngc, local_gc_start)."""
    ngc = [local_n[t] + gc_below[t] + gc_above[t] for t in range(len(local_n))]
    gcs = [local_start[t] - gc_below[t] for t in range(len(local_n))]
    mem = howmany
    for x in ngc:
        mem *= x
    return mem, ngc, gcs


# ---- a11: test-data contract (api/api-basic.c:60-121,148-189,254-276) -----------
def _c_mod(a, n):
    """C's % (truncation toward zero) on numpy int arrays."""
    return np.fmod(a, n)


def _plain_index(n, kvecs):
    """api/api-basic.c:254-264: k += k*n[t] + kvec[t]  =>  Horner with (n[t]+1)."""
    k = np.zeros_like(kvecs[0])
    for t in range(len(n)):
        k = k + k * n[t] + kvecs[t]
    return k


def _init_scalar(n, gvecs):
    per = [_c_mod(gvecs[t], n[t]) for t in range(len(n))]
    m = _plain_index(n, per).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        val = 1000.0 / (2 * m) + 1j * (1000.0 / (2 * m + 1))
    return np.where(m == 0, 1500.0 + 1250.0j, val)


def init_input(kind, n, local_n, local_start):
    """kind in {'complex','complex_hermitian','real'}; returns the local block in
    row-major order with shape local_n (api/api-basic.c:88-121)."""
    grids = np.meshgrid(*[np.arange(local_n[t], dtype=np.int64) + local_start[t]
                          for t in range(len(n))], indexing="ij")
    d1 = _init_scalar(n, grids)
    if kind == "complex":
        return d1
    if kind == "real":
        return np.where(grids[-1] < n[-1], d1.real, 0.0)
    mirrored = [n[t] - grids[t] for t in range(len(n))]
    d2 = _init_scalar(n, mirrored)
    return 0.5 * (d1 + np.conj(d2))


def check_output(kind, n, local_n, local_start, data):
    """max |data - pattern| over the local block (api/api-basic.c:148-189)."""
    ref = init_input(kind, n, local_n, local_start)
    data = np.asarray(data).reshape(ref.shape)
    if kind == "real":
        grids = np.meshgrid(*[np.arange(local_n[t], dtype=np.int64) + local_start[t]
                              for t in range(len(n))], indexing="ij")
        err = np.where(grids[-1] < n[-1], np.abs(data - ref), 0.0)
    else:
        err = np.abs(data - ref)
    return float(err.max()) if err.size else 0.0


# ---- transforms on the gathered (global) array ------------------------------------
def _pad_axis(x, axis, n_to, zl):
    pad = [(0, 0)] * x.ndim
    pad[axis] = (zl, n_to - x.shape[axis] - zl)
    return np.pad(x, pad)


def _cut_axis(x, axis, zl, d):
    sl = [slice(None)] * x.ndim
    sl[axis] = slice(zl, zl + d)
    return x[tuple(sl)]


def _sign_vec(length, start, half, extra):
    """(-1)^g for g < half else 1 (api/api-basic.c:1213-1234), g = pos + start."""
    g = np.arange(length, dtype=np.int64) + start
    f = np.where(g < half, np.where(g % 2 != 0, -1.0, 1.0) * extra, 1.0)
    return f


def _r2r_1d(x, kind, axis):
    import scipy.fft as sf
    table = {REDFT00: ("dct", 1), REDFT10: ("dct", 2), REDFT01: ("dct", 3), REDFT11: ("dct", 4),
             RODFT00: ("dst", 1), RODFT10: ("dst", 2), RODFT01: ("dst", 3), RODFT11: ("dst", 4)}
    if kind not in table:
        raise NotImplementedError("r2r kind %d (halfcomplex/DHT) is out of scope" % kind)
    fn, typ = table[kind]
    return getattr(sf, fn)(x, type=typ, axis=axis, norm=None)


def global_transform(kind, x, n, ni=None, no=None, sign=FORWARD, flags=0, kinds=None,
                     skip=None, howmany=1):
    """What one pfft_execute computes, expressed on the whole array.

    x: global input, shape ni (+ (howmany,) if howmany > 1); for r2c the logical
       real array (no padding), for c2r the complex array with last dim ni/2+1.
    Returns the global output with shape no (c2c/r2r), (.., no/2+1) (r2c) or no
    (c2r, real, unpadded).

    Follows execute_full (api/api-basic.c:1044-1107): [conj] -> twiddle_input
    (SHIFTED_OUT) -> per dim: embed ni->n (kernel/ousample.c:208-339), 1-D
    transform (kernel/sertrafo.c:518-544 = FFTW's unnormalised DFT), trunc
    n->no -> twiddle_output (SHIFTED_IN) -> [conj].
    """
    d = len(n)
    ni = list(n) if ni is None else list(ni)
    no = list(n) if no is None else list(no)
    skip = [0] * d if skip is None else list(skip)
    x = np.asarray(x)
    cdt = np.complex128
    si, so = bool(flags & SHIFTED_IN), bool(flags & SHIFTED_OUT)

    if kind == C2R and sign == FORWARD:   # kernel/partrafo.c:428-443
        x = np.conj(x)
    work = x.astype(np.float64 if kind in (R2C, R2R) else cdt)

    # twiddle_input (api/api-basic.c:1186-1238)
    if so:
        for t in range(d):
            if skip[t]:
                continue
            start = -(ni[t] // 2) if si else 0
            extra = -1.0 if (si and (n[t] // 2) % 2) else 1.0
            f = _sign_vec(work.shape[t], start, ni[t] // 2, extra)
            shp = [1] * work.ndim
            shp[t] = work.shape[t]
            work = work * f.reshape(shp)

    order = list(range(d - 1, -1, -1))  # last dim first, as the forward schedule does
    if kind == C2R:
        order = list(range(d))          # c2r transforms on the way back: dim 0 first, real dim last
    for t in order:
        last = t == d - 1
        # ---- embed ni[t] -> n[t]
        if kind == C2R and last:
            pni, pn = ni[t] // 2 + 1, n[t] // 2 + 1
            work = _pad_axis(work, t, pn, pn - pni)          # Zl = pno - pni (ousample.c:292-301)
        else:
            zl = (n[t] - ni[t]) // 2 if si else 0            # ousample.c:247-258
            work = _pad_axis(work, t, n[t], zl)
        # ---- transform
        if not skip[t]:
            if kind == R2R:
                work = _r2r_1d(work, kinds[t], t)
            elif kind == R2C and last:
                work = np.fft.rfft(work, axis=t)             # forward; BACKWARD handled by final conj
            elif kind == C2R and last:
                work = np.fft.irfft(work, n=n[t], axis=t) * n[t]
            else:
                s = sign
                if kind == R2C:
                    s = FORWARD
                if kind == C2R:
                    s = BACKWARD
                work = np.fft.fft(work, axis=t) if s == FORWARD else np.fft.ifft(work, axis=t) * n[t]
        elif kind == R2C and last:
            raise NotImplementedError("skipping the r2c dimension")
        # ---- trunc n[t] -> no[t]
        if kind == R2C and last:
            pn, pno = n[t] // 2 + 1, no[t] // 2 + 1
            work = _cut_axis(work, t, pn - pno, pno)         # Zl = pni - pno (ousample.c:262-272)
        else:
            zl = (n[t] - no[t]) // 2 if so else 0
            work = _cut_axis(work, t, zl, no[t])

    # twiddle_output (api/api-basic.c:1242-1285)
    if si:
        for t in range(d):
            if skip[t]:
                continue
            start = -(no[t] // 2) if so else 0
            f = _sign_vec(work.shape[t], start, no[t] // 2, 1.0)
            shp = [1] * work.ndim
            shp[t] = work.shape[t]
            work = work * f.reshape(shp)

    if kind == R2C and sign == BACKWARD:
        work = np.conj(work)
    return work


def brute_force_dft(x, sign=FORWARD):
    """O(n^2)-per-axis DFT straight from the definition (doc/features.tex:140-156);
    independent of pocketfft, used to pin global_transform on small sizes."""
    x = np.asarray(x, dtype=np.complex128)
    for ax in range(x.ndim):
        nn = x.shape[ax]
        j = np.arange(nn)
        w = np.exp(sign * 2j * np.pi * np.outer(j, j) / nn)
        x = np.moveaxis(np.tensordot(w, x, axes=([1], [ax])), 0, ax)
    return x


# ---- scatter / gather between the global array and a rank's local block -------------
def mem_order(rnk_n, rnk_pm, transposed):
    """Memory order of a local block: natural, or dims (1..r, 0, r+1..) for the
    transposed layout (doc/tutorial.tex:322-325; api/api-basic.c:1149-1183)."""
    if not transposed:
        return list(range(rnk_n))
    if rnk_pm >= rnk_n:      # 3-D data on a 3-D mesh runs on a 2-D mesh internally (kernel/procmesh.c:191-205)
        rnk_pm = rnk_n - 1
    return list(range(1, rnk_pm + 1)) + [0] + list(range(rnk_pm + 1, rnk_n))


def extract_block(glob, local_n, local_start, order=None, shift=None):
    """Local block of `glob` as a C-contiguous array in memory order `order`."""
    d = len(local_n)
    shift = [0] * d if shift is None else shift
    sl = tuple(slice(local_start[t] + shift[t], local_start[t] + shift[t] + local_n[t]) for t in range(d))
    blk = glob[sl]
    if order is not None:
        blk = np.transpose(blk, order + list(range(d, blk.ndim)))
    return np.ascontiguousarray(blk)


def place_block(glob, block, local_n, local_start, order=None, shift=None):
    d = len(local_n)
    shift = [0] * d if shift is None else shift
    sl = tuple(slice(local_start[t] + shift[t], local_start[t] + shift[t] + local_n[t]) for t in range(d))
    if order is not None:
        shp = [local_n[t] for t in order] + list(glob.shape[d:])
        block = np.asarray(block).reshape(shp)
        inv = np.argsort(order).tolist()
        block = np.transpose(block, inv + list(range(d, block.ndim)))
    else:
        block = np.asarray(block).reshape([local_n[t] for t in range(d)] + list(glob.shape[d:]))
    glob[sl] = block


# ---- a12: ghost cells (gcell/gcells_plan.c, gcells_sendrecv.c) ------------
def gc_exchange_block(glob, n, local_n, local_start, gc_below, gc_above):
    """Net effect of pfft_exchange on one rank: dense block of shape
    gc_below+local_n+gc_above holding the global array at indices
    local_start-gc_below ... taken mod n (periodic)."""
    idx = [np.mod(np.arange(local_start[t] - gc_below[t], local_start[t] + local_n[t] + gc_above[t]), n[t])
           for t in range(len(n))]
    return glob[np.ix_(*idx)] if glob.ndim == len(n) else glob[np.ix_(*idx)]


def gc_reduce_global(n, blocks):
    """Net effect of pfft_reduce: blocks = [(gc_block, local_n, local_start,
    gc_below, gc_above)] for every rank; returns the global array in which every
    owned element is the sum of itself and all its halo copies."""
    first = blocks[0][0]
    glob = np.zeros(list(n) + list(first.shape[len(n):]), dtype=first.dtype)
    for blk, ln, ls, gb, ga in blocks:
        idx = [np.mod(np.arange(ls[t] - gb[t], ls[t] + ln[t] + ga[t]), n[t]) for t in range(len(n))]
        np.add.at(glob, np.ix_(*idx), blk)
    return glob


# ---- synthetic benchmark input ------------------------------------------
def splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def synthetic_complex(n, local_n, local_start, seed=1234):
    """Uniform [-1,1) re/im from a counter-based hash of the GLOBAL linear index,
    so every mesh sees the same global array."""
    grids = np.meshgrid(*[np.arange(local_n[t], dtype=np.uint64) + np.uint64(local_start[t])
                          for t in range(len(n))], indexing="ij")
    lin = np.zeros_like(grids[0])
    for t in range(len(n)):
        lin = lin * np.uint64(n[t]) + grids[t]
    with np.errstate(over="ignore"):
        re = splitmix64(np.uint64(seed) ^ (lin * np.uint64(2)))
        im = splitmix64(np.uint64(seed) ^ (lin * np.uint64(2) + np.uint64(1)))
    to_f = lambda u: (u >> np.uint64(11)).astype(np.float64) * (2.0 ** -52) - 1.0
    return to_f(re) + 1j * to_f(im)

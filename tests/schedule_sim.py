"""tests/schedule_sim.py -- TEST INFRASTRUCTURE ONLY.

Executes, in numpy and for all ranks of a virtual mesh at once, the stage lists the
planner of pfft_b200 emits (`pfftb200_describe_schedule`, include/pfft_b200.h).  It
interprets exactly the fields the CUDA stage kernel interprets (strides, chunking,
embed/truncate windows, +-1 modulations), so a planner bug shows up on the CPU box;
the per-line transform itself is numpy's.  Never imported by the product.
"""
import ctypes as C
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INT = C.c_ssize_t
KIND = {"c2c": 0, "r2c": 1, "c2r": 2, "r2r": 3}
OP_COPY, OP_C2C, OP_R2C, OP_C2R, OP_R2R = range(5)


def load_lib():
    from pfft_b200 import _lib
    return _lib.load()


def _vec(v):
    return None if v is None else (INT * len(v))(*v)


def describe(lib, kind, n, np_, pid, flags=0, ni=None, no=None, howmany=1, iblock=None, oblock=None,
             sign=-1, kinds=None, skip=None):
    d = len(n)
    ni = n if ni is None else ni
    no = n if no is None else no
    fn = lib.pfftb200_describe_schedule
    fn.restype = C.c_size_t
    args = [C.c_int(KIND[kind]), C.c_int(d), _vec(n), _vec(ni), _vec(no), INT(howmany), _vec(iblock), _vec(oblock),
            C.c_int(len(np_)), (C.c_int * len(np_))(*np_), C.c_int(pid), C.c_int(sign),
            None if kinds is None else (C.c_int * d)(*kinds),
            None if skip is None else (C.c_int * (len(np_) + 1))(*skip), C.c_uint(flags)]
    need = fn(*args, None, C.c_size_t(0))
    buf = C.create_string_buffer(need + 16)
    fn(*args, buf, C.c_size_t(need + 16))
    return json.loads(buf.value.decode())


def _sign_mod(mod, length):
    on, start, half, extra = mod
    if not on:
        return None
    g = np.arange(length, dtype=np.int64) + start
    return np.where(g < half, np.where(g % 2 != 0, -1.0, 1.0) * extra, 1.0)


def _r2r(line, kind):
    import scipy.fft as sf
    table = {3: ("dct", 1), 5: ("dct", 2), 4: ("dct", 3), 6: ("dct", 4),
             7: ("dst", 1), 9: ("dst", 2), 8: ("dst", 3), 10: ("dst", 4)}
    fn, typ = table[kind]
    return getattr(sf, fn)(line, type=typ, axis=1, norm=None)


def run_stage(g, inbuf):
    """Returns the list of output chunks (one per output segment) of one stage on one rank."""
    nb = len(g["batch"])
    lin = np.zeros((), dtype=np.int64)
    lout = np.zeros((), dtype=np.int64)
    swz_c = np.zeros((), dtype=np.int64)     # batch coordinate that drives the producer-side swizzle
    for k, (ext, is_, os_, _dim) in enumerate(g["batch"]):
        idx = np.arange(ext, dtype=np.int64)
        shape = [1] * nb
        shape[k] = ext
        lin = lin + (idx * is_).reshape(shape)
        lout = lout + (idx * os_).reshape(shape)
        if k == g.get("oswz_batch", -1):
            swz_c = swz_c + idx.reshape(shape)
    swz_c = np.broadcast_to(swz_c, np.broadcast(lin, lout, swz_c).shape).reshape(-1) if nb else np.zeros(1, dtype=np.int64)
    lin = np.asarray(lin).reshape(-1)
    lout = np.asarray(lout).reshape(-1)
    if any(b[0] == 0 for b in g["batch"]):
        lin = lin[:0]
        lout = lout[:0]
    line_id = None
    if g.get("ntile", 0) > 0:
        # explicit tiles (micro-blocked layouts): batch[] enumerates tiles, the tables place the lines of a tile
        nt = g["ntile"]
        tio = np.asarray(g["tile_ioff"], dtype=np.int64)
        too = np.asarray(g["tile_ooff"], dtype=np.int64)
        l = np.arange(nt, dtype=np.int64)
        m = (swz_c & g.get("oswz_mask", 0)) << g.get("oswz_shift", 0)
        lout = (lout[:, None] + (too[None, :] ^ m[:, None])).reshape(-1)
        if g.get("iswz_mask", 0):
            line_id = np.tile(l, lin.size)           # the input offset of a line depends on the block index
            lin = np.repeat(lin, nt)
        else:
            lin = (lin[:, None] + tio[None, :]).reshape(-1)
    out_dtype = np.float64 if g["out_real"] else np.complex128
    chunks = [np.zeros(c, dtype=out_dtype) for c in g["oseg_cnt"]]
    nin, n, nout = g["nin"], g["n"], g["nout"]
    if lin.size == 0 or nout == 0:
        return chunks
    j = np.arange(nin, dtype=np.int64)
    seg = j // g["iblk"]
    jl = j - seg * g["iblk"]
    b2 = g.get("iblk2", 1)
    in_idx = seg * g["iseg_stride"] + (jl * g["istride"] if b2 == 1 else (jl // b2) * g["iblk2_stride"] + (jl % b2) * g["istride"])
    if line_id is None:
        X = inbuf[lin[:, None] + in_idx[None, :]]
    else:
        blk = ((jl // b2) & g["iswz_mask"]) * b2     # XOR on the in-block offset, just above the line-index part
        tio = np.asarray(g["tile_ioff"], dtype=np.int64)
        X = inbuf[lin[:, None] + (tio[line_id][:, None] ^ blk[None, :]) + in_idx[None, :]]
    if g["conj_in"]:
        X = np.conj(X)
    f = _sign_mod(g["mod_in"], nin)
    if f is not None:
        X = X * f[None, :]
    op = g["op"]
    if op == OP_C2R:
        length = n // 2 + 1
    else:
        length = n
    line = np.zeros((lin.size, length), dtype=X.dtype)
    line[:, g["zin"]:g["zin"] + nin] = X
    if op == OP_COPY:
        Y = line
    elif op == OP_C2C:
        Y = np.fft.fft(line, axis=1) if g["sign"] < 0 else np.fft.ifft(line, axis=1) * n
    elif op == OP_R2C:
        Y = np.fft.rfft(line.real, n=n, axis=1)
    elif op == OP_C2R:
        Y = np.fft.irfft(line, n=n, axis=1) * n
    else:
        Y = _r2r(line.real, g["r2r_kind"])
    Y = Y[:, g["zout"]:g["zout"] + nout]
    f = _sign_mod(g["mod_out"], nout)
    if f is not None:
        Y = Y * f[None, :]
    if g["conj_out"]:
        Y = np.conj(Y)
    if g["out_real"]:
        Y = Y.real
    k = np.arange(nout, dtype=np.int64)
    kseg = k // g["oblk"]
    for q in range(g["noseg"]):
        ks = k[kseg == q]
        if ks.size == 0:
            continue
        kl = ks - q * g["oblk"]
        b2 = g.get("oblk2", 1)
        pos = lout[:, None] + (kl * g["ostride"] if b2 == 1 else (kl // b2) * g["oblk2_stride"] + (kl % b2) * g["ostride"])[None, :]
        chunks[q][pos] = Y[:, ks]
    return chunks


def rank_of(np_, coords):
    r = 0
    for t in range(len(np_)):
        r = r * np_[t] + coords[t]
    return r


def simulate(scheds, user_in, stage_fn=None):
    """scheds: list over ranks of schedule dicts; user_in: list over ranks of flat input arrays.
    Returns list over ranks of flat output arrays (the user's out buffers).
    stage_fn(rank, stage index, stage dict, input buffer) -> list of output chunks replaces the numpy
    interpretation of a stage (tests/test_kernel_emulation.py plugs in the emulated CUDA kernel body)."""
    P = len(scheds)
    nst = len(scheds[0]["stages"])
    assert all(len(s["stages"]) == nst for s in scheds)
    cur = list(user_in)
    for i in range(nst):
        if stage_fn is None:
            outs = [run_stage(scheds[r]["stages"][i], cur[r]) for r in range(P)]
        else:
            outs = [stage_fn(r, i, scheds[r]["stages"][i], cur[r]) for r in range(P)]
        xi = scheds[0]["stages"][i]["exchange"]
        if xi < 0:
            cur = [np.concatenate(o) if len(o) > 1 else o[0] for o in outs]
            continue
        nxt = []
        for r in range(P):
            x = scheds[r]["exchanges"][xi]
            dtype = np.float64 if x["elem_real"] else np.complex128
            nxt.append(np.full(x["recv_cnt"] * x["nparts"], np.nan, dtype=dtype))
        for r in range(P):
            s = scheds[r]
            x = s["exchanges"][xi]
            grp = s["groups"][x["mesh_dim"]]
            assert grp["size"] == x["nparts"] and grp["me"] == x["me"] and grp["members"][x["me"]] == r
            for q in range(x["nparts"]):
                dst = grp["members"][q]
                xd = scheds[dst]["exchanges"][xi]
                chunk = outs[r][q]
                assert chunk.size == x["send_cnt"][q]
                assert chunk.size == xd["recv_cnt"], (chunk.size, xd["recv_cnt"])
                nxt[dst][x["me"] * xd["recv_cnt"]:(x["me"] + 1) * xd["recv_cnt"]] = chunk
        cur = nxt
    return cur

"""The BODY of the any-length CUDA stage kernel (pfft_b200/csrc/fft_mixed.h -- the source stage_mixed_kernel
executes on the GPU) run on the CPU, threads emulated between barriers (pfftb200_emulate_stage), for every
stage of every rank of a virtual mesh; exchanges are done in numpy by tests/schedule_sim.py.  The gathered
result must equal the oracle.  This checks the kernel's arithmetic (codelets, Stockham passes, packed real
lines, Bluestein, DCT/DST twiddles) and its addressing (chunks, windows, modulations) without a GPU; the
-m gpu tests then check the same code as compiled for sm_100a."""
import ctypes as C
import os

import numpy as np
import pytest

import cases
import pfft_b200 as pf
import pfft_oracle as po
import schedule_sim as ss

T_IN, T_OUT, PAD = po.TRANSPOSED_IN, po.TRANSPOSED_OUT, po.PADDED_R2C
S_IN, S_OUT = po.SHIFTED_IN, po.SHIFTED_OUT
KIND = {"c2c": 0, "r2c": 1, "c2r": 2, "r2r": 3}


def emulated_stage_fn(case, prec):
    lib = pf.lib()
    fn = lib.pfftb200_emulate_stage
    fn.restype = C.c_int
    n = case["n"]
    d = len(n)
    np_ = case["np"]
    INT = pf.INT
    rdt, cdt = (np.float64, np.complex128) if prec == 0 else (np.float32, np.complex64)

    def vec(v, t=INT):
        return None if v is None else (t * len(v))(*[int(x) for x in v])

    def run(rank, i, g, inbuf):
        src = np.ascontiguousarray(np.nan_to_num(np.asarray(inbuf), nan=777.0).astype(rdt if g["in_real"] else cdt))
        outs = [np.zeros(max(c, 1), dtype=rdt if g["out_real"] else cdt) for c in g["oseg_cnt"]]
        ptrs = (C.c_void_p * len(outs))(*[o.ctypes.data for o in outs])
        rc = fn(C.c_int(prec), C.c_int(KIND[case["kind"]]), C.c_int(d), vec(n), vec(case.get("ni") or n), vec(case.get("no") or n),
                INT(case.get("howmany", 1)), vec(case.get("iblock")), vec(case.get("oblock")), C.c_int(len(np_)), vec(np_, C.c_int),
                C.c_int(rank), C.c_int(case.get("sign", -1)), vec(case.get("kinds"), C.c_int), vec(case.get("skip"), C.c_int),
                C.c_uint(case.get("flags", 0)), C.c_int(i), C.c_void_p(src.ctypes.data), ptrs)
        assert rc == 0, pf.last_error()
        dt = np.float64 if g["out_real"] else np.complex128
        return [o[:c].astype(dt) for o, c in zip(outs, g["oseg_cnt"])]

    return run


def run_emulated(case, prec=0):
    os.environ["PFFT_B200_BLOCKED"] = "0"      # plain layouts: the micro-blocked ones belong to the register-resident kernel
    try:
        np_ = case["np"]
        P = int(np.prod(np_))
        r = len(np_)
        scheds = [pf.describe_schedule(case["kind"], case["n"], np_, pid, case.get("flags", 0), case.get("ni"), case.get("no"),
                                       case.get("howmany", 1), case.get("iblock"), case.get("oblock"), case.get("sign", -1),
                                       case.get("kinds"), case.get("skip")) for pid in range(P)]
        for s in scheds:
            assert s["error"] == "", s["error"]
        xg = cases.make_global_input(case, 0)
        user_in = [cases.local_input(case, xg, s["local_ni"], s["local_i_start"], r) for s in scheds]
        outs = ss.simulate(scheds, user_in, stage_fn=emulated_stage_fn(case, prec))
        want = cases.oracle_output(case, xg)
        scale = max(1e-300, float(np.abs(want).max()))
        err = max(cases.compare_local_output(case, want, outs[pid], s["local_no"], s["local_o_start"], r)
                  for pid, s in enumerate(scheds))
        return err / scale
    finally:
        os.environ.pop("PFFT_B200_BLOCKED", None)


CASES = [
    dict(kind="c2c", n=[29, 27, 31], np=[2, 2]),                       # BASELINE config 1: 27 = 3^3, 29 and 31 through Bluestein
    dict(kind="c2c", n=[29, 27, 31], np=[1, 1], flags=T_OUT),
    dict(kind="c2c", n=[29, 27, 31], np=[2, 2], flags=T_IN, sign=+1),
    dict(kind="c2c", n=[8, 6, 4], np=[1, 1]),
    dict(kind="c2c", n=[16, 12, 10], np=[4]),
    dict(kind="c2c", n=[13, 14, 19, 17], np=[2, 2, 2], flags=T_OUT),
    dict(kind="c2c", n=[5, 4, 3], np=[3, 2]),
    dict(kind="c2c", n=[8, 6, 4], np=[2, 2], howmany=3),
    dict(kind="c2c", n=[16, 12, 10], np=[2, 2], iblock=[9, 7], oblock=[10, 8]),
    dict(kind="c2c", n=[8, 6, 4], np=[2, 2], skip=[0, 1, 0]),
    dict(kind="c2c", n=[96, 48, 80], np=[2, 2], flags=T_OUT),          # 3 * 2^k, 5 * 2^k
    dict(kind="c2c", n=[60, 35, 77], np=[1, 2], flags=T_OUT),          # radices 3, 4, 5, 7, 11
    dict(kind="c2c", n=[26, 39, 64], np=[2, 1], flags=T_IN, sign=+1),  # radix 13, 16 * 4
    dict(kind="c2c", n=[128, 8, 256], np=[1, 1]),                      # 16 * 8, 16 * 16
    dict(kind="c2c", n=[32, 512, 4], np=[1, 1], flags=T_OUT),          # 8 * 4, 16 * 8 * 4
    dict(kind="c2c", n=[4, 4, 202], np=[1, 1]),                        # 2 * 101: Bluestein on a composite length
    dict(kind="r2c", n=[29, 27, 31], np=[2, 2], flags=T_OUT),          # odd real lines: full-length complex
    dict(kind="c2r", n=[29, 27, 31], np=[2, 2], flags=T_IN, sign=+1),
    dict(kind="r2c", n=[16, 12, 10], np=[2, 2], flags=T_OUT | PAD),    # even real lines: n/2 packed points
    dict(kind="c2r", n=[16, 12, 10], np=[2, 2], flags=T_IN | PAD, sign=+1),
    dict(kind="r2c", n=[16, 12, 10], np=[2, 2], flags=T_OUT, sign=+1),
    dict(kind="c2r", n=[16, 12, 10], np=[2, 2], flags=T_IN, sign=-1),
    dict(kind="r2c", n=[8, 6, 62], np=[1, 1]),                         # packed points through Bluestein (31)
    dict(kind="c2r", n=[8, 6, 62], np=[1, 1], sign=+1),
    dict(kind="r2c", n=[8, 16, 128], np=[2, 2], flags=T_OUT),
    dict(kind="c2r", n=[8, 16, 128], np=[2, 2], flags=T_IN, sign=+1),
    dict(kind="r2c", n=[4, 6, 768], np=[1, 1], ni=[4, 4, 512], no=[4, 6, 768]),   # config 5's oversampled real line
    dict(kind="c2r", n=[4, 6, 768], np=[1, 1], ni=[4, 6, 768], no=[4, 4, 512], sign=+1),
    dict(kind="c2c", n=[12, 10, 9], ni=[6, 5, 4], no=[12, 10, 9], np=[2, 2]),
    dict(kind="c2c", n=[12, 10, 9], ni=[12, 10, 9], no=[5, 7, 3], np=[2, 2], flags=T_OUT),
    dict(kind="r2c", n=[29, 27, 31], ni=[16, 16, 16], no=[29, 27, 31], np=[2, 2], flags=T_OUT),
    dict(kind="c2r", n=[29, 27, 31], ni=[29, 27, 31], no=[16, 16, 16], np=[2, 2], flags=T_IN, sign=+1),
    dict(kind="r2c", n=[12, 10, 16], ni=[6, 5, 10], no=[12, 10, 16], np=[2, 2], flags=T_OUT),   # pruned packed lines
    dict(kind="c2r", n=[12, 10, 16], ni=[12, 10, 16], no=[6, 5, 10], np=[2, 2], flags=T_IN, sign=+1),
    dict(kind="c2c", n=[8, 6, 4], np=[2, 2], flags=S_IN | S_OUT),
    dict(kind="c2c", n=[16, 12, 8], ni=[8, 6, 4], no=[16, 12, 8], np=[2, 2], flags=S_IN | S_OUT),
    dict(kind="c2r", n=[8, 16, 128], np=[2, 2], flags=S_OUT, sign=+1),
    dict(kind="c2r", n=[8, 16, 128], np=[1, 1], flags=S_IN | S_OUT | T_IN, sign=+1),
    dict(kind="r2c", n=[8, 16, 128], np=[2, 2], flags=S_IN | S_OUT | T_OUT),
    dict(kind="c2c", n=[29, 27, 31], np=[2, 2, 2]),
    dict(kind="r2c", n=[29, 27, 31], np=[2, 2, 2], flags=T_OUT),
    dict(kind="c2r", n=[29, 27, 31], np=[2, 2, 2], sign=+1),
    dict(kind="r2r", n=[13, 11, 9], np=[2, 2], kinds=[po.REDFT00, po.REDFT01, po.REDFT10]),
    dict(kind="r2r", n=[13, 11, 9], np=[2, 2], kinds=[po.RODFT00, po.RODFT10, po.REDFT11], flags=T_OUT),
    dict(kind="r2r", n=[13, 11, 9], np=[2, 2], kinds=[po.RODFT01, po.RODFT11, po.REDFT00], flags=T_IN),
    dict(kind="r2r", n=[12, 10, 9], ni=[6, 5, 4], no=[12, 10, 9], np=[2, 2], kinds=[po.REDFT00, po.RODFT00, po.REDFT01]),
    dict(kind="r2r", n=[64, 32, 128], np=[1, 1], kinds=[po.REDFT10, po.RODFT10, po.REDFT01], flags=T_OUT),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: "%s-%s-np%s-f%d%s" % (
    c["kind"], "x".join(map(str, c["n"])), "x".join(map(str, c["np"])), c.get("flags", 0), "-pruned" if "ni" in c else ""))
def test_emulated_kernel_body_reproduces_oracle(built_lib, case):
    assert run_emulated(case, 0) < 1e-12


@pytest.mark.parametrize("case", [CASES[0], CASES[10], CASES[18], CASES[19], CASES[22], CASES[26], CASES[-1]],
                         ids=lambda c: "%s-%s" % (c["kind"], "x".join(map(str, c["n"]))))
def test_emulated_kernel_body_single_precision(built_lib, case):
    assert run_emulated(case, 1) < 2e-5


def test_long_lines(built_lib):
    """Lengths beyond what shared memory holds (the GPU uses a global workspace for them; the arithmetic is
    the same): 12288 = 3 * 4096 and the prime 8191 (Bluestein, M = 16384)."""
    for n in ([2, 2, 12288], [2, 2, 8191]):
        assert run_emulated(dict(kind="c2c", n=n, np=[1, 1]), 0) < 1e-12

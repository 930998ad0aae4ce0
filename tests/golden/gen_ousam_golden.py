"""tests/golden/gen_ousam_golden.py -- regenerates tests/golden/ousam_embed_trunc.json.

Runs ONLY in the build container (needs /root/reference): calls the reference's own embed / truncate
code (kernel/ousample.c: pfft_plan_ousam_dd, pfft_execute_ousam_dd, compiled by `make -C oracle` into
oracle/_ref/libpfft_refint.so) on index-valued rows and records where every input element lands.
These are the conventions of SURVEY.md 8 (a8): zeros at the end, half/half with PFFT_SHIFTED_*, the
padding of r2c input / c2r output rows, and the quirk that truncated r2c output / embedded c2r input
keep the UPPER end of the half spectrum (kernel/ousample.c:262-272,292-301).

    make -C oracle && python tests/golden/gen_ousam_golden.py
"""
import ctypes as C
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import refint  # noqa: E402

INT = C.c_ssize_t
C2C, R2C, C2R, PADDED = 1 << 0, 1 << 1, 1 << 2, 1 << 13      # kernel/ipfft.h:138-151
EMBED, TRUNC = 1 << 0, 1 << 1                                 # kernel/ipfft.h:167-168
S_IN, S_OUT = 1 << 2, 1 << 3                                  # api/pfft.h:530-531
SENTINEL = -7.0


def run_case(lib, trafo, ousam, si, n0, n1i, n1o, hm):
    """One 1-D embed / truncate of n0 rows: returns (reals per input row, reals per output row, out list)."""
    lib.pfft_local_size_ousam_dd.restype = INT
    lib.pfft_plan_ousam_dd.restype = C.c_void_p
    ni, no = (INT * 1)(n1i), (INT * 1)(n1o)
    mem = lib.pfft_local_size_ousam_dd(INT(n0), C.c_int(1), ni, no, INT(hm), C.c_uint(trafo))
    cnt = 2 * mem + 8
    Buf = C.c_double * cnt
    a, b = Buf(), Buf()
    for k in range(cnt):
        a[k] = k + 1.0
        b[k] = SENTINEL
    plan = lib.pfft_plan_ousam_dd(INT(n0), C.c_int(1), ni, no, INT(hm), a, b, C.c_uint(trafo), C.c_uint(si), C.c_uint(ousam))
    assert plan, "reference refused the parameters"
    lib.pfft_execute_ousam_dd(C.c_void_p(plan), a, b, a, b)
    return [b[k] for k in range(cnt)]


def main():
    ref = refint.RefInt()
    lib = ref.lib
    cases = []
    for trafo, name in ((C2C, "c2c"), (R2C, "r2c"), (R2C | PADDED, "r2c_padded"), (C2R, "c2r"), (C2R | PADDED, "c2r_padded")):
        for ousam, oname in ((EMBED, "embed"), (TRUNC, "trunc")):
            for si in (0, S_IN | S_OUT):
                for n0, small, big, hm in ((1, 4, 8, 1), (3, 6, 10, 1), (2, 5, 9, 2), (2, 8, 12, 1)):
                    if si and (small % 2 or big % 2):
                        continue
                    n1i, n1o = (small, big) if ousam == EMBED else (big, small)
                    out = run_case(lib, trafo, ousam, si, n0, n1i, n1o, hm)
                    cases.append(dict(trafo=name, op=oname, shifted=bool(si), n0=n0, n1i=n1i, n1o=n1o, howmany=hm, out=out))
    path = os.path.join(HERE, "ousam_embed_trunc.json")
    json.dump(cases, open(path, "w"))
    print("wrote %d cases to %s" % (len(cases), path))


if __name__ == "__main__":
    main()

"""tests/golden/gen_shift_golden.py -- regenerates tests/golden/shift_twiddles.json.

Runs ONLY in the build container (needs /root/reference): calls the reference's own twiddle_input /
twiddle_output (api/api-basic.c:1186-1285, static -- reached through oracle/stubs/expose_static.c) on an
array of ones covering the WHOLE index range of a one-rank problem and records the +-1 factor of every
element: the conventions behind PFFT_SHIFTED_IN / PFFT_SHIFTED_OUT.

    make -C oracle && python tests/golden/gen_shift_golden.py
"""
import ctypes as C
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import refint  # noqa: E402

INT = C.c_ssize_t
C2C = 1 << 0                                     # PFFTI_TRAFO_C2C, kernel/ipfft.h:138
T_IN, T_OUT, S_IN, S_OUT = 1, 2, 4, 8            # api/pfft.h:528-531


def factors(lib, output_side, n, nio, skip, transposed, pfft_flags, rnk_pm=2):
    """Factor of every element of the one-rank block, in MEMORY order (transposed layouts: dims 1..r, 0, rest)."""
    d = len(n)
    shifted = pfft_flags & (S_OUT if output_side else S_IN)
    start = [-(v // 2) if shifted else 0 for v in nio]       # local_*_start of a one-rank block (kernel/partrafo.c:178-190)
    total = 1
    for v in nio:
        total *= v
    Buf = C.c_double * (2 * total)
    a, b = Buf(*([1.0] * (2 * total))), Buf()
    V = INT * d
    transp = (T_OUT if output_side else T_IN) if transposed else 0
    lib.oracle_ref_twiddle(C.c_int(output_side), C.c_int(d), C.c_int(rnk_pm), V(*n), V(*nio), V(*nio), V(*start),
                           (C.c_int * (rnk_pm + 1))(*skip), INT(1), C.c_uint(C2C), C.c_uint(transp), C.c_uint(pfft_flags), a, b)
    out = [int(b[2 * k]) for k in range(total)]
    assert all(b[2 * k + 1] == b[2 * k] for k in range(total)) and set(out) <= {1, -1}
    return out


def main():
    lib = refint.RefInt().lib
    cases = []
    for n, nio in (([8, 6, 4], [8, 6, 4]), ([8, 6, 4], [4, 6, 2]), ([4, 10, 6], [4, 6, 6]), ([12, 4, 8], [6, 2, 8])):
        for flags in (S_IN, S_OUT, S_IN | S_OUT):
            for skip in ([0, 0, 0], [0, 1, 0]):
                for transposed in (False, True):
                    for output_side in (0, 1):
                        # twiddle_input exists for SHIFTED_OUT plans, twiddle_output for SHIFTED_IN plans (api/api-basic.c:1063-1096)
                        if not (flags & (S_IN if output_side else S_OUT)):
                            continue
                        cases.append(dict(output_side=output_side, n=n, nio=nio, flags=flags, skip=skip, transposed=transposed,
                                          factors=factors(lib, output_side, n, nio, skip, transposed, flags)))
    path = os.path.join(HERE, "shift_twiddles.json")
    json.dump(cases, open(path, "w"))
    print("wrote %d cases to %s" % (len(cases), path))


if __name__ == "__main__":
    main()

"""CPU tests: the oracle (oracle/pfft_oracle.py) against the golden vectors captured
from the reference's own compiled integer code (tests/golden/gen_golden.py), against
the values quoted in SURVEY.md 8c, and -- when the prebuilt oracle/_ref library is
present -- against that library live."""
import json
import os

import numpy as np
import pytest

import pfft_oracle as po


def _load(golden_dir, name):
    with open(os.path.join(golden_dir, name)) as f:
        return json.load(f)


def test_local_block_matches_reference_golden(golden_dir):
    g = _load(golden_dir, "local_block.json")
    assert len(g["cases"]) > 10000
    for kind, ni, no, np_, pid, flags, ib, ob, lni, lis, lno, los in g["cases"]:
        got = po.local_block(kind, ni, no, np_, pid, flags, ib, ob)
        assert [list(x) for x in got] == [lni, lis, lno, los], (kind, ni, no, np_, pid, flags, ib, ob)


def test_local_size_gc_matches_reference_golden(golden_dir):
    g = _load(golden_dir, "local_size_gc.json")
    for c in g["cases"]:
        mem, ngc, gcs = po.local_size_gc(c["local_n"], c["local_start"], c["howmany"], c["gc_below"], c["gc_above"])
        assert (mem, ngc, gcs) == (c["mem"], c["ngc"], c["gc_start"])


def test_init_input_bit_exact(golden_dir):
    g = _load(golden_dir, "init_input.json")
    for c in g["cases"]:
        got = po.init_input(c["kind"], c["n"], c["local_n"], c["local_start"]).ravel()
        if c["kind"] == "real":
            want = np.array([float.fromhex(v) for v in c["values"]])
        else:
            want = np.array([complex(float.fromhex(a), float.fromhex(b)) for a, b in c["values"]])
        assert got.shape == want.shape
        assert np.array_equal(got.view(np.float64), want.view(np.float64)), c["kind"]


def test_survey_quoted_values():
    # values printed by the reference's own code
    x = po.init_input("complex", [29, 27, 31], [2, 2, 3], [0, 0, 0]).ravel()
    assert x[0] == 1500 + 1250j and x[2] == 250 + 200j
    assert x[3] == complex(15.625, 15.384615384615385)
    assert po.local_block("c2c", [29, 27, 31], [29, 27, 31], [2, 2], 3, po.TRANSPOSED_OUT) == \
        ([14, 13, 31], [15, 14, 0], [29, 13, 15], [0, 14, 16])
    assert po.local_block("r2c", [1024] * 3, [1024] * 3, [2, 4], 7, po.TRANSPOSED_OUT | po.PADDED_R2C) == \
        ([512, 256, 1026], [512, 768, 0], [1024, 512, 126], [0, 512, 387])
    assert po.local_block("c2c", [29, 27, 31], [29, 27, 31], [2, 2, 2], 7, po.TRANSPOSED_OUT)[2:] == \
        ([29, 13, 7], [0, 14, 24])
    assert po.local_block("c2c", [128] * 4, [128] * 4, [2, 2, 2], 7, po.TRANSPOSED_OUT) == \
        ([64, 64, 64, 128], [64, 64, 64, 0], [128, 64, 64, 64], [0, 64, 64, 64])


def test_live_reference_library_if_present():
    import refint
    if not refint.available():
        pytest.skip("oracle/_ref/libpfft_refint.so not built (needs /root/reference)")
    ref = refint.RefInt()
    rng = np.random.default_rng(7)
    for _ in range(300):
        d = int(rng.integers(3, 5))
        n = [int(x) for x in rng.integers(2, 40, size=d)]
        mesh = [int(x) for x in rng.integers(1, 4, size=int(rng.integers(1, 3)))]
        kind = ["c2c", "r2c", "c2r", "r2r"][int(rng.integers(0, 4))]
        tf = int(rng.integers(0, 3))
        if (kind == "r2c" and tf == 1) or (kind == "c2r" and tf == 2):
            tf = 0
        P = int(np.prod(mesh))
        pid = int(rng.integers(0, P))
        name = {"c2c": "dft", "r2c": "dft_r2c", "c2r": "dft_c2r", "r2r": "r2r"}[kind]
        want = ref.local_block(name, n, n, mesh, pid, tf)
        got = po.local_block(kind, n, n, mesh, pid, tf)
        assert tuple(map(list, want)) == tuple(map(list, got))


def test_transform_oracle_against_dft_definition():
    """pocketfft restatement vs the O(n^2) definition (forward values are not pinned
    by any reference test; the DFT definition is the anchor)."""
    rng = np.random.default_rng(3)
    x = rng.standard_normal((5, 7, 6)) + 1j * rng.standard_normal((5, 7, 6))
    for sign in (po.FORWARD, po.BACKWARD):
        a = po.global_transform(po.C2C, x, [5, 7, 6], sign=sign)
        b = po.brute_force_dft(x, sign)
        assert np.linalg.norm(a - b) / np.linalg.norm(b) < 1e-13
    xr = rng.standard_normal((4, 5, 9))
    a = po.global_transform(po.R2C, xr, [4, 5, 9])
    b = po.brute_force_dft(xr, po.FORWARD)[:, :, :5]
    assert np.linalg.norm(a - b) / np.linalg.norm(b) < 1e-13
    back = po.global_transform(po.C2R, a, [4, 5, 9], sign=po.BACKWARD)
    assert np.allclose(back, xr * xr.size, atol=1e-10)


def test_survey_numpy_forward_values():
    # SURVEY.md 8c (derived by the restated oracle): forward of the a11 pattern at 29x27x31
    n = [29, 27, 31]
    x = po.init_input("complex", n, n, [0, 0, 0])
    X = po.global_transform(po.C2C, x, n)
    assert abs(X[0, 0, 0] - (6687.614919969376 + 6131.20324008727j)) < 1e-8
    assert abs(X[1, 2, 3] - (2198.945614340035 + 739.7122735812995j)) < 1e-8
    assert abs(np.linalg.norm(X) - 3.288979778416490e+05) < 1e-6


def test_pruned_and_roundtrip():
    rng = np.random.default_rng(5)
    ni, n, no = [4, 6, 5], [8, 9, 7], [8, 9, 7]
    x = rng.standard_normal(ni) + 1j * rng.standard_normal(ni)
    X = po.global_transform(po.C2C, x, n, ni=ni, no=no)
    xp = np.zeros(n, complex)
    xp[:4, :6, :5] = x
    assert np.allclose(X, np.fft.fftn(xp))
    back = po.global_transform(po.C2C, X, n, ni=no, no=ni, sign=po.BACKWARD)
    assert np.allclose(back, x * np.prod(n))


def test_shifted_semantics():
    """SHIFTED_IN|SHIFTED_OUT means: user index g in [-n/2, n/2) on both sides,
    Y[k] = sum_j X[j] exp(-2 pi i j k / n) with j,k centred (doc/features.tex)."""
    n = [8, 6, 4]
    rng = np.random.default_rng(11)
    x = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    Y = po.global_transform(po.C2C, x, n, flags=po.SHIFTED_IN | po.SHIFTED_OUT)
    want = np.fft.fftshift(np.fft.fftn(np.fft.ifftshift(x)))
    assert np.allclose(Y, want)


# ---- embed / truncate conventions pinned on the reference's own kernel/ousample.c ----------
def _stage_of_dim(kind, n, ni, no, flags, dim):
    import pfft_b200 as pf
    s = pf.describe_schedule(kind, n, [1, 1], 0, flags, ni, no, 1, None, None, +1 if kind == "c2r" else -1)
    assert s["error"] == "", s["error"]
    st = [g for g in s["stages"] if g["dim"] == dim and g["op"] != 0]
    assert len(st) == 1
    return st[0]


def test_embed_truncate_windows_match_reference_code(built_lib, golden_dir):
    """Every row the reference's embed / truncate loops produce (tests/golden/gen_ousam_golden.py) is what
    the product planner's stage windows (nin, zin) / (nout, zout) describe -- including the padding of real
    rows and the upper-end quirk of truncated r2c output and embedded c2r input."""
    cases = _load(golden_dir, "ousam_embed_trunc.json")
    assert len(cases) >= 60
    S_IN, S_OUT, PAD = po.SHIFTED_IN, po.SHIFTED_OUT, po.PADDED_R2C
    for c in cases:
        trafo, op, hm, n0, n1i, n1o = c["trafo"], c["op"], c["howmany"], c["n0"], c["n1i"], c["n1o"]
        kind = trafo.split("_")[0]
        padded = trafo.endswith("padded")
        embed = op == "embed"
        n = [2, 2, n1o if embed else n1i]
        ni = [2, 2, n1i] if embed else n
        no = n if embed else [2, 2, n1o]
        flags = (S_IN | S_OUT if c["shifted"] else 0) | (PAD if padded else 0)
        g = _stage_of_dim(kind, n, ni, no, flags, 2)
        pn = lambda v: v // 2 + 1     # noqa: E731  physical (complex) length of a real line's spectrum
        # element width in reals and row lengths in elements on both sides of the reference's 1-D operation
        if kind == "c2c":
            w, row_i, row_o = 2 * hm, n1i, n1o
        elif kind == "r2c" and embed:          # real input rows; the output rows carry the r2c padding
            w, row_i, row_o = hm, (2 * pn(n1i) if padded else n1i), 2 * pn(n1o)
        elif kind == "r2c":                    # half spectra
            w, row_i, row_o = 2 * hm, pn(n1i), pn(n1o)
        elif embed:                            # c2r, half spectra
            w, row_i, row_o = 2 * hm, pn(n1i), pn(n1o)
        else:                                  # c2r output rows: padding removed unless PADDED
            w, row_i, row_o = hm, 2 * pn(n1i), (2 * pn(n1o) if padded else n1o)
        if embed:
            cnt, off = g["nin"], g["zin"]
        else:
            cnt, off = g["nout"], g["zout"]
        src = np.arange(1, n0 * row_i * w + 1, dtype=np.float64).reshape(n0, row_i, w)
        want = np.full((n0, row_o, w), np.nan)
        if embed:
            total = n1o if kind != "c2r" else pn(n1o)          # the length-n line the window sits in
            want[:, :total, :] = 0.0
            want[:, off:off + cnt, :] = src[:, :cnt, :]
        else:
            want[:, :cnt, :] = src[:, off:off + cnt, :]
        got = np.asarray(c["out"][:want.size]).reshape(want.shape)
        written = ~np.isnan(want)
        assert np.array_equal(got[written], want[written]), (c["trafo"], op, c["shifted"], n0, n1i, n1o, hm, g["nin"], g["zin"], g["nout"], g["zout"])
        assert np.all(got[~written] == -7.0)       # the reference leaves the padding of real rows untouched


def test_embed_truncate_golden_is_current_if_reference_present(golden_dir):
    import refint
    if not (refint.available() and os.path.isdir("/root/reference")):
        pytest.skip("oracle/_ref not built here")
    import importlib.util
    spec = importlib.util.spec_from_file_location("gen_ousam", os.path.join(golden_dir, "gen_ousam_golden.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    lib = refint.RefInt().lib
    c = _load(golden_dir, "ousam_embed_trunc.json")[5]
    trafo = {"c2c": gen.C2C, "r2c": gen.R2C, "r2c_padded": gen.R2C | gen.PADDED, "c2r": gen.C2R, "c2r_padded": gen.C2R | gen.PADDED}[c["trafo"]]
    out = gen.run_case(lib, trafo, gen.EMBED if c["op"] == "embed" else gen.TRUNC, (gen.S_IN | gen.S_OUT) if c["shifted"] else 0,
                       c["n0"], c["n1i"], c["n1o"], c["howmany"])
    assert out == c["out"]


# ---- PFFT_SHIFTED_IN / _OUT modulations pinned on the reference's own twiddle code ----------
def test_shift_modulations_match_reference_code(built_lib, golden_dir):
    """The +-1 factor of EVERY element as the reference's twiddle_input / twiddle_output compute it
    (tests/golden/gen_shift_golden.py) equals the product of the per-dimension modulations the planner fuses
    into the stage that transforms that dimension (Stage::mod_in / mod_out) -- pruned sizes, skipped
    transforms, both-sides-shifted extra sign and transposed memory order included."""
    import pfft_b200 as pf
    cases = _load(golden_dir, "shift_twiddles.json")
    assert len(cases) >= 60
    for c in cases:
        n, nio, out_side = c["n"], c["nio"], c["output_side"]
        tflag = (po.TRANSPOSED_OUT if out_side else po.TRANSPOSED_IN) if c["transposed"] else 0
        ni, no = (n, nio) if out_side else (nio, n)
        s = pf.describe_schedule("c2c", n, [1, 1], 0, c["flags"] | tflag, ni, no, 1, None, None, -1, None, c["skip"])
        assert s["error"] == "", s["error"]
        vecs = []
        for t in range(3):
            f = np.ones(nio[t])
            for g in s["stages"]:
                if g["dim"] != t or g["op"] == 0:
                    continue
                on, start, half, extra = g["mod_out"] if out_side else g["mod_in"]
                assert (g["nout"] if out_side else g["nin"]) == nio[t]
                if on:
                    idx = np.arange(nio[t]) + start
                    f = np.where(idx < half, np.where(idx % 2 != 0, -1.0, 1.0) * extra, 1.0)
            vecs.append(f)
        full = vecs[0][:, None, None] * vecs[1][None, :, None] * vecs[2][None, None, :]
        if c["transposed"]:
            full = full.transpose(1, 2, 0)          # memory order of a transposed layout on a 2-D mesh
        want = np.asarray(c["factors"], dtype=np.float64).reshape(full.shape)
        assert np.array_equal(full, want), {k: v for k, v in c.items() if k != "factors"}

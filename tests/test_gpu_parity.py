"""GPU parity tests (-m gpu): the CUDA path behind the C ABI against the oracle.

fp64 bar: relative max error <= 1e-12 (north_star: rel. L2 <= 1e-12); fp32: <= 1e-5.
Small sizes go through the numpy oracle on the gathered array; the reference's own
round-trip check (pfft_init_input_* -> forward -> clear -> backward -> scale ->
pfft_check_output_*, tests/simple_check_c2c.c:17-69, tol 1e-12 in tests/run_checks.sh:75)
is run at small and at BASELINE sizes."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import pfft_oracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

T_IN, T_OUT, PAD = po.TRANSPOSED_IN, po.TRANSPOSED_OUT, po.PADDED_R2C
S_IN, S_OUT = po.SHIFTED_IN, po.SHIFTED_OUT
TOL = {"double": 1e-12, "single": 2e-5}        # relative MAX error (stricter than the stated metric in fp64)
TOL_L2 = {"double": 1e-12, "single": 1e-5}     # north_star: relative L2 error of the whole array


@pytest.fixture(scope="module")
def world(built_lib):
    import pfft_b200 as pf
    pf.init()
    return pf


def _id(c):
    return "%s-%s-np%s-f%d-%s%s" % (c["kind"], "x".join(map(str, c["n"])), "x".join(map(str, c["np"])),
                                    c.get("flags", 0), c.get("precision", "double")[0], c.get("tag", ""))


SINGLE_RANK = [
    dict(kind="c2c", n=[8, 8, 8], np=[1, 1]),
    dict(kind="c2c", n=[8, 8, 8], np=[1, 1], sign=+1),
    dict(kind="c2c", n=[29, 27, 31], np=[1, 1]),                 # the reference test's odd sizes
    dict(kind="c2c", n=[29, 27, 31], np=[1, 1], flags=T_OUT),
    dict(kind="c2c", n=[29, 27, 31], np=[1, 1], flags=T_IN, sign=+1),
    dict(kind="c2c", n=[64, 32, 128], np=[1, 1], flags=T_OUT),
    dict(kind="c2c", n=[64, 32, 128], np=[1, 1], flags=T_IN, sign=+1),
    dict(kind="c2c", n=[64, 64, 64], np=[1]),
    dict(kind="c2c", n=[128, 16, 256], np=[1, 1]),
    dict(kind="c2c", n=[16, 8, 1024], np=[1, 1], flags=T_OUT),
    dict(kind="c2c", n=[1024, 8, 16], np=[1, 1], flags=T_OUT),
    dict(kind="c2c", n=[12, 10, 18, 6], np=[1, 1, 1], flags=T_OUT),
    dict(kind="c2c", n=[16, 16, 16, 16], np=[1, 1, 1]),
    dict(kind="c2c", n=[32, 20], np=[1]),
    dict(kind="c2c", n=[8, 6, 4], np=[1, 1], howmany=3),
    dict(kind="c2c", n=[64, 32, 16], np=[1, 1], precision="single"),
    dict(kind="c2c", n=[29, 27, 31], np=[1, 1], precision="single", flags=T_OUT),
    dict(kind="c2c", n=[16, 16, 16], np=[1, 1], memory="host"),
    dict(kind="c2c", n=[16, 16, 16], np=[1, 1], inplace=True),
    dict(kind="c2c", n=[16, 16, 16], np=[1, 1], inplace=True, flags=T_OUT),
    dict(kind="r2c", n=[16, 12, 10], np=[1, 1]),
    dict(kind="r2c", n=[29, 27, 31], np=[1, 1], flags=T_OUT),
    dict(kind="r2c", n=[32, 16, 64], np=[1, 1], flags=T_OUT | PAD),
    dict(kind="r2c", n=[16, 12, 10], np=[1, 1], sign=+1),
    dict(kind="c2r", n=[16, 12, 10], np=[1, 1], sign=+1),
    dict(kind="c2r", n=[29, 27, 31], np=[1, 1], flags=T_IN, sign=+1),
    dict(kind="c2r", n=[32, 16, 64], np=[1, 1], flags=T_IN | PAD, sign=+1),
    dict(kind="c2r", n=[16, 12, 10], np=[1, 1], sign=-1),
    dict(kind="r2c", n=[32, 16, 64], np=[1, 1], precision="single", flags=T_OUT | PAD),
    # real lines of power-of-two length through the register-resident kernel (full-length complex transform)
    dict(kind="r2c", n=[8, 16, 256], np=[1, 1]),
    dict(kind="c2r", n=[8, 16, 256], np=[1, 1], sign=+1),
    dict(kind="r2c", n=[16, 8, 1024], np=[1, 1], flags=T_OUT | PAD, precision="single"),
    dict(kind="c2r", n=[16, 8, 1024], np=[1, 1], flags=T_IN | PAD, sign=+1, precision="single"),
    dict(kind="c2r", n=[8, 16, 128], np=[1, 1], flags=T_IN, sign=-1),
    dict(kind="r2c", n=[8, 16, 128], np=[1, 1], flags=T_OUT, sign=+1),
    dict(kind="r2c", n=[12, 10, 128], ni=[6, 5, 64], no=[12, 10, 128], np=[1, 1], flags=T_OUT),
    dict(kind="c2c", n=[12, 10, 9], ni=[6, 5, 4], no=[12, 10, 9], np=[1, 1]),
    dict(kind="c2c", n=[12, 10, 9], ni=[12, 10, 9], no=[5, 7, 3], np=[1, 1], flags=T_OUT),
    dict(kind="r2c", n=[29, 27, 31], ni=[16, 16, 16], no=[29, 27, 31], np=[1, 1], flags=T_OUT),
    dict(kind="c2r", n=[29, 27, 31], ni=[29, 27, 31], no=[16, 16, 16], np=[1, 1], flags=T_IN, sign=+1),
    dict(kind="c2c", n=[8, 6, 4], np=[1, 1], flags=S_IN | S_OUT),
    dict(kind="c2c", n=[8, 6, 4], np=[1, 1], flags=S_IN | S_OUT | T_OUT),
    dict(kind="c2c", n=[16, 12, 8], ni=[8, 6, 4], no=[16, 12, 8], np=[1, 1], flags=S_IN | S_OUT),
    dict(kind="c2c", n=[8, 6, 4], np=[1, 1], skip=[0, 1, 0]),
    # index shifts on real transforms whose last dimension takes the register-resident kernels
    dict(kind="c2r", n=[8, 16, 128], np=[1, 1], flags=S_OUT, sign=+1),
    dict(kind="c2r", n=[8, 16, 128], np=[1, 1], flags=S_IN | S_OUT, sign=+1),
    dict(kind="r2c", n=[8, 16, 128], np=[1, 1], flags=S_IN),
    dict(kind="r2c", n=[8, 16, 128], np=[1, 1], flags=S_IN | S_OUT | T_OUT),
    dict(kind="c2c", n=[8, 16, 128], np=[1, 1], flags=S_IN | S_OUT),
    # r2r: DCT/DST kinds (FFTW enum values; reference tests/simple_check_r2r.c uses REDFT00/01/10, RODFT00)
    dict(kind="r2r", n=[9, 8, 7], np=[1, 1], kinds=[po.REDFT00, po.REDFT01, po.REDFT10]),
    dict(kind="r2r", n=[9, 8, 7], np=[1, 1], kinds=[po.RODFT00, po.RODFT01, po.RODFT10], flags=T_OUT),
    dict(kind="r2r", n=[6, 5, 12], np=[1, 1], kinds=[po.REDFT11, po.RODFT11, po.REDFT10], flags=T_IN),
    dict(kind="r2r", n=[16, 16, 16], np=[1, 1], kinds=[po.REDFT10, po.REDFT10, po.REDFT10], precision="single"),
    dict(kind="r2r", n=[12, 10, 9], ni=[6, 5, 4], no=[12, 10, 9], np=[1, 1], kinds=[po.REDFT00, po.RODFT00, po.REDFT01]),
    dict(kind="r2r", n=[64, 32, 128], np=[1, 1], kinds=[po.REDFT10, po.RODFT10, po.REDFT01], flags=T_OUT),
    # register-resident kernel for real lines (packed half-length transforms) and lengths 3 * 2^k (fft_reg_kernel.h)
    dict(kind="r2c", n=[8, 4, 2048], np=[1, 1], flags=T_OUT, tag="-reg"),
    dict(kind="c2r", n=[8, 4, 2048], np=[1, 1], flags=T_IN, sign=+1, tag="-reg"),
    dict(kind="r2c", n=[4, 8, 4096], np=[1, 1], tag="-reg"),
    dict(kind="c2r", n=[4, 8, 4096], np=[1, 1], sign=+1, tag="-reg"),
    dict(kind="r2c", n=[16, 8, 512], np=[1, 1], flags=T_OUT | PAD, tag="-reg"),
    dict(kind="c2r", n=[16, 8, 512], np=[1, 1], flags=T_IN | PAD, sign=+1, tag="-reg"),
    dict(kind="r2c", n=[8, 8, 1024], np=[1, 1], flags=T_OUT, tag="-reg"),
    dict(kind="c2r", n=[8, 8, 1024], np=[1, 1], flags=T_IN, sign=+1, precision="single", tag="-reg"),
    dict(kind="r2c", n=[8, 16, 256], ni=[8, 16, 128], no=[8, 16, 256], np=[1, 1], flags=T_OUT, tag="-reg"),
    dict(kind="c2r", n=[8, 16, 256], ni=[8, 16, 256], no=[8, 16, 128], np=[1, 1], flags=T_IN, sign=+1, tag="-reg"),
    dict(kind="r2c", n=[8, 16, 256], ni=[8, 16, 128], no=[8, 16, 256], np=[1, 1], flags=S_IN | S_OUT, tag="-reg"),
    dict(kind="r2c", n=[8, 16, 256], ni=[8, 16, 256], no=[4, 8, 128], np=[1, 1], flags=T_OUT, tag="-reg"),      # truncated half spectrum (upper end kept)
    dict(kind="c2r", n=[8, 16, 256], ni=[4, 8, 128], no=[8, 16, 256], np=[1, 1], flags=T_IN, sign=+1, tag="-reg"),
    dict(kind="r2c", n=[8, 16, 768], ni=[8, 16, 768], no=[4, 8, 384], np=[1, 1], tag="-reg"),
    dict(kind="c2r", n=[8, 16, 768], ni=[4, 8, 384], no=[8, 16, 768], np=[1, 1], sign=+1, precision="single", tag="-reg"),
    dict(kind="c2c", n=[4, 8, 768], np=[1, 1], tag="-reg"),
    dict(kind="c2c", n=[4, 8, 768], np=[1, 1], sign=+1, flags=T_OUT, tag="-reg"),
    dict(kind="c2c", n=[768, 4, 8], np=[1, 1], flags=T_OUT, tag="-reg"),
    dict(kind="c2c", n=[8, 384, 16], np=[1, 1], flags=T_OUT, tag="-reg"),
    dict(kind="c2c", n=[8, 16, 192], np=[1, 1], flags=T_IN, sign=+1, tag="-reg"),
    dict(kind="c2c", n=[4, 4, 1536], np=[1, 1], flags=T_OUT, precision="single", tag="-reg"),
    dict(kind="c2c", n=[8, 768, 12], ni=[8, 512, 8], no=[8, 768, 12], np=[1, 1], flags=T_OUT, tag="-reg"),
    dict(kind="c2c", n=[8, 768, 12], ni=[8, 768, 12], no=[8, 512, 8], np=[1, 1], flags=T_IN, sign=+1, tag="-reg"),
    dict(kind="c2c", n=[8, 4, 768], ni=[8, 4, 512], no=[8, 4, 768], np=[1, 1], flags=S_IN | S_OUT, tag="-reg"),
    dict(kind="r2c", n=[8, 4, 768], np=[1, 1], flags=T_OUT, tag="-reg"),
    dict(kind="c2r", n=[8, 4, 768], np=[1, 1], flags=T_IN, sign=+1, tag="-reg"),
    dict(kind="r2c", n=[4, 8, 384], np=[1, 1], flags=T_OUT | PAD, precision="single", tag="-reg"),
    dict(kind="c2r", n=[4, 8, 384], np=[1, 1], flags=T_IN | PAD, sign=+1, precision="single", tag="-reg"),
    dict(kind="r2c", n=[4, 4, 1536], np=[1, 1], tag="-reg"),
    dict(kind="c2r", n=[4, 4, 3072], np=[1, 1], sign=+1, tag="-reg"),
    dict(kind="r2c", n=[12, 12, 768], ni=[8, 8, 512], no=[12, 12, 768], np=[1, 1], flags=T_OUT, tag="-reg"),       # config 5 in small
    dict(kind="c2r", n=[12, 12, 768], ni=[12, 12, 768], no=[8, 8, 512], np=[1, 1], flags=T_IN, sign=+1, tag="-reg"),
    # long lines of the any-length kernel: large prime (Bluestein through the global workspace), 3 * 4096, prime 6397
    dict(kind="c2c", n=[4, 4, 8191], np=[1, 1], tag="-long"),
    dict(kind="c2c", n=[4, 12288, 4], np=[1, 1], flags=T_OUT, tag="-long"),
    dict(kind="c2c", n=[8, 8, 6397], np=[1, 1], tag="-long"),
    dict(kind="r2c", n=[4, 4, 8192], np=[1, 1], precision="single", tag="-long"),
    # plane-fused last pair (power-of-two lines of equal length in the last two stages)
    dict(kind="c2c", n=[64, 64, 64], np=[1, 1], flags=T_OUT, tag="-fused"),
    dict(kind="c2c", n=[64, 64, 64], np=[1, 1], flags=T_IN, sign=+1, tag="-fused"),
    dict(kind="c2c", n=[128, 128, 128], np=[1, 1], flags=T_OUT, tag="-fused"),
    dict(kind="c2c", n=[128, 128, 128], np=[1, 1], flags=T_IN, sign=+1, tag="-fused"),
    dict(kind="c2c", n=[128, 128, 128], np=[1, 1], flags=T_OUT, inplace=True, tag="-fused"),
    dict(kind="c2c", n=[128, 128, 64], np=[1, 1], flags=T_OUT, tag="-fused"),
    dict(kind="c2c", n=[64, 128, 128], np=[1, 1], flags=T_IN, sign=+1, tag="-fused"),
    dict(kind="c2c", n=[128, 128, 128], np=[1, 1], flags=T_OUT, precision="single", tag="-fused"),
    dict(kind="c2c", n=[64, 512, 512], np=[1, 1], flags=T_IN, sign=+1, tag="-fused"),
    dict(kind="c2c", n=[512, 512, 32], np=[1, 1], flags=T_OUT, repeat=2, tag="-fused"),
    dict(kind="c2c", n=[64, 64, 64, 64], np=[1, 1, 1], flags=T_OUT, tag="-fused"),
]


@pytest.mark.parametrize("case", SINGLE_RANK, ids=_id)
def test_single_rank_matches_oracle(world, case):
    import gpu_worker
    comm = world.create_procmesh(case["np"])
    res = gpu_worker.run_case(case, comm)
    comm.free()
    assert res["error"] == "", res["error"]
    err, l2 = gpu_worker.check_case(case, [res], l2=True)
    assert err < TOL[case.get("precision", "double")], err
    assert l2 < TOL_L2[case.get("precision", "double")], l2
    assert res["input_preserved"], "out-of-place plans must not touch the input (PFFT_PRESERVE_INPUT default)"
    if case.get("tag") == "-reg":
        assert "reg" in res["kernels"], res["kernels"]
    if case.get("tag") == "-fused" and case.get("precision", "double") == "double" and os.environ.get("PFFT_B200_FUSE", "0") == "1":
        assert res["fused"] == 1, "expected the plane-fused pair kernel for this case"


MULTI_RANK = [
    dict(kind="c2c", n=[29, 27, 31], np=[2, 2]),                       # BASELINE configs[0]
    dict(kind="c2c", n=[29, 27, 31], np=[2, 2], flags=T_OUT),
    dict(kind="c2c", n=[29, 27, 31], np=[2, 2], flags=T_IN, sign=+1),
    dict(kind="c2c", n=[64, 32, 128], np=[2, 2], flags=T_OUT),
    dict(kind="c2c", n=[16, 12, 10], np=[4]),
    dict(kind="c2c", n=[13, 14, 19, 17], np=[2, 2, 2], flags=T_OUT),
    dict(kind="c2c", n=[5, 4, 3], np=[3, 2]),
    dict(kind="r2c", n=[29, 27, 31], np=[2, 2], flags=T_OUT),
    dict(kind="c2r", n=[29, 27, 31], np=[2, 2], flags=T_IN, sign=+1),
    dict(kind="c2c", n=[32, 32, 32], np=[2, 2], flags=T_OUT, precision="single"),
    dict(kind="r2c", n=[29, 27, 31], ni=[16, 16, 16], no=[29, 27, 31], np=[2, 2], flags=T_OUT),
    # 3-D data on a 3-D mesh (3dto2d remap)
    dict(kind="c2c", n=[29, 27, 31], np=[2, 2, 2]),
    dict(kind="c2c", n=[29, 27, 31], np=[2, 2, 2], flags=T_OUT),
    dict(kind="c2c", n=[29, 27, 31], np=[2, 2, 2], flags=T_IN, sign=+1),
    dict(kind="r2c", n=[29, 27, 31], np=[2, 2, 2], flags=T_OUT),
    dict(kind="c2c", n=[64, 64, 64], np=[1, 2, 4], flags=T_OUT),
    dict(kind="r2c", n=[64, 64, 128], np=[2, 2], flags=T_OUT, tag="-pow2real"),
    dict(kind="c2r", n=[64, 64, 128], np=[2, 2], flags=T_IN, sign=+1, tag="-pow2real"),
    dict(kind="c2r", n=[8, 16, 128], np=[2, 2], flags=S_OUT, sign=+1),
    dict(kind="c2r", n=[8, 16, 128], np=[2, 2], flags=S_IN | S_OUT | T_IN, sign=+1),
    dict(kind="r2c", n=[8, 16, 128], np=[2, 2], flags=S_IN | S_OUT | T_OUT),
    # register-resident real lines / 3 * 2^k lengths across exchanges (segmented outputs, gathered inputs)
    dict(kind="r2c", n=[16, 16, 512], np=[2, 2], flags=T_OUT, tag="-reg"),
    dict(kind="c2r", n=[16, 16, 512], np=[2, 2], flags=T_IN, sign=+1, tag="-reg"),
    dict(kind="r2c", n=[8, 768, 768], ni=[8, 512, 512], no=[8, 768, 768], np=[2, 2], flags=T_OUT, tag="-reg"),
    dict(kind="c2r", n=[8, 768, 768], ni=[8, 768, 768], no=[8, 512, 512], np=[2, 2], flags=T_IN, sign=+1, tag="-reg"),
    dict(kind="c2c", n=[384, 12, 192], np=[2, 2], tag="-reg"),
    dict(kind="c2c", n=[384, 12, 192], np=[3], flags=T_OUT, precision="single", tag="-reg"),
    # micro-blocked chains of power-of-two stages across exchanges
    dict(kind="c2c", n=[64, 64, 64], np=[2, 2], flags=T_OUT, tag="-blk"),
    dict(kind="c2c", n=[64, 64, 64], np=[2, 2], flags=T_IN, sign=+1, tag="-blk"),
    dict(kind="c2c", n=[128, 128, 128], np=[2, 4], flags=T_OUT, tag="-blk"),
    dict(kind="c2c", n=[64, 64, 64], np=[4, 1], flags=T_OUT, precision="single", tag="-blk"),
    dict(kind="r2r", n=[13, 11, 9], np=[2, 2], kinds=[po.REDFT00, po.REDFT01, po.REDFT10]),
    dict(kind="r2r", n=[13, 11, 9], np=[2, 2], kinds=[po.RODFT00, po.RODFT10, po.REDFT11], flags=T_OUT),
    dict(kind="r2r", n=[7, 6, 5, 4], np=[2, 2, 2], kinds=[po.REDFT10, po.RODFT00, po.REDFT00, po.RODFT01], flags=T_OUT),
]


def run_multi(case, timeout=300, env_extra=None):
    import gpu_worker
    P = int(np.prod(case["np"]))
    with tempfile.TemporaryDirectory() as td:
        json.dump(case, open(os.path.join(td, "case.json"), "w"))
        env = dict(os.environ)
        env.update(env_extra or {})
        cmd = [os.path.join(ROOT, "pfft_b200", "bin", "pfftrun"), "-np", str(P), "-timeout", str(timeout),
               sys.executable, os.path.join(ROOT, "tests", "gpu_worker.py"), os.path.join(td, "case.json"), td]
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout + 30, env=env)
        assert p.returncode == 0, (p.returncode, p.stdout[-2000:], p.stderr[-2000:])
        results = []
        for rk in range(P):
            meta = json.load(open(os.path.join(td, "rank%d.json" % rk)))
            assert meta.get("error", "") == "", meta["error"]
            meta["out"] = np.load(os.path.join(td, "rank%d.npy" % rk))
            results.append(meta)
    return gpu_worker.check_case(case, results, l2=True), results


@pytest.mark.parametrize("case", MULTI_RANK, ids=_id)
def test_multi_rank_on_one_gpu_matches_oracle(built_lib, case):
    """N ranks, rank % device_count -> GPU (they share GPU 0 on a one-GPU box); exchanges go through
    peer-mapped (CUDA IPC) stores.  The one-rank-per-GPU runs are in tests/test_gpu_multi.py."""
    (err, l2), results = run_multi(case)
    assert err < TOL[case.get("precision", "double")], err
    assert l2 < TOL_L2[case.get("precision", "double")], l2
    assert all(r["transport"] == "p2p" for r in results)


def _round_trip(world, n, np_, flags_f, flags_b, dtype=np.complex128, device_memory=False):
    """The reference's own check (tests/simple_check_c2c.c): returns pfft_check_output's max error."""
    pf = world
    comm = pf.create_procmesh(np_)
    alloc, lni, lis, lno, los = pf.local_size("c2c", n, comm, flags_f)
    a = pf.alloc_complex(alloc, dtype)
    b = pf.alloc_complex(alloc, dtype)
    fwd = pf.plan_dft(n, a, b, comm, pf.FORWARD, flags_f | pf.DESTROY_INPUT, dtype=dtype)
    bwd = pf.plan_dft(n, b, a, comm, pf.BACKWARD, flags_b | pf.DESTROY_INPUT, dtype=dtype)
    assert fwd is not None and bwd is not None, pf.last_error()
    pf.init_input("complex", n, lni, lis, a, dtype=dtype)
    fwd.execute()
    pf.clear_input("complex", n, lni, lis, a, dtype=dtype)
    bwd.execute()
    cnt = int(np.prod(lni))
    a.array[:cnt] /= float(np.prod(n))
    err = pf.check_output("complex", n, lni, lis, a, comm, dtype=dtype)
    fwd.destroy(); bwd.destroy(); a.free(); b.free(); comm.free()
    return err


def test_reference_round_trip_small(world):
    assert _round_trip(world, [29, 27, 31], [1, 1], 0, 0) < 1e-12
    assert _round_trip(world, [29, 27, 31], [1, 1], T_OUT, T_IN) < 1e-12
    assert _round_trip(world, [16, 16, 16], [1, 1], T_OUT, T_IN, dtype=np.complex64) < 1e-4


def test_reference_round_trip_baseline_size(world):
    """BASELINE configs[1] shape on one GPU: 512^3 c2c fp64, TRANSPOSED_OUT / TRANSPOSED_IN."""
    err = _round_trip(world, [512, 512, 512], [1, 1], T_OUT, T_IN)
    assert err < 1e-12, err


@pytest.mark.parametrize("n,dtype,tol", [([512, 512, 512], "float64", 1e-12), ([512, 1024, 512], "float64", 1e-12),
                                         ([256, 256, 256], "float64", 1e-12), ([512, 512, 512], "float32", 2e-5)])
def test_forward_values_at_baseline_size(world, n, dtype, tol):
    """BASELINE-size forward output, element by element, against a library FFT of the same array
    (torch.fft = cuFFT, used here ONLY as a second checker where the numpy oracle takes minutes);
    TRANSPOSED_OUT layout [k1][k2][k0].  Covers the 8-line micro-blocked tiles of the headline size."""
    import torch
    pf = world
    rdt = getattr(torch, dtype)
    cdt = torch.complex128 if dtype == "float64" else torch.complex64
    comm = pf.create_procmesh([1, 1])
    N = int(np.prod(n))
    g = torch.Generator(device="cuda").manual_seed(7)
    x = torch.rand(N, 2, dtype=rdt, device="cuda", generator=g) * 2 - 1
    X = torch.empty_like(x)
    npdt = np.complex128 if dtype == "float64" else np.complex64
    plan = pf.plan_dft(n, x, X, comm, pf.FORWARD, T_OUT, dtype=npdt)
    assert plan is not None, pf.last_error()
    plan.execute(x, X)
    want = torch.fft.fftn(torch.view_as_complex(x).reshape(n)).permute(1, 2, 0).contiguous()
    got = torch.view_as_complex(X).reshape(n[1], n[2], n[0])
    err = ((got - want).abs().max() / want.abs().max()).item()
    del want
    # and back through TRANSPOSED_IN
    y = torch.empty_like(x)
    back = pf.plan_dft(n, X, y, comm, pf.BACKWARD, T_IN, dtype=npdt)
    back.execute(X, y)
    err_rt = ((y / N - x).abs().max()).item()
    plan.destroy(); back.destroy(); comm.free()
    assert err < tol and err_rt < 10 * tol, (err, err_rt)
    assert all(k == "pow2" for k in plan_kernels(pf, n))


def plan_kernels(pf, n):
    s = pf.describe_schedule("c2c", n, [1, 1], 0, T_OUT)
    return ["pow2" if g["ntile"] > 0 else "plain" for g in s["stages"]]


def test_linearity_and_parseval_at_size(world):
    """Size-independent properties at 256^3: Parseval and linearity of the forward transform."""
    import torch
    pf = world
    n = [256, 256, 256]
    comm = pf.create_procmesh([1, 1])
    N = int(np.prod(n))
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(N, 2, dtype=torch.float64, device="cuda", generator=g)
    y = torch.randn(N, 2, dtype=torch.float64, device="cuda", generator=g)
    X = torch.empty_like(x); Y = torch.empty_like(x); Z = torch.empty_like(x)
    plan = pf.plan_dft(n, x, X, comm, pf.FORWARD, T_OUT)
    plan.execute(x, X)
    plan.execute(y, Y)
    z = 2.0 * x - 3.0 * y
    plan.execute(z, Z)
    lin = (Z - (2.0 * X - 3.0 * Y)).norm() / Z.norm()
    pars = abs((X.norm() ** 2 / N - x.norm() ** 2) / x.norm() ** 2)
    plan.destroy(); comm.free()
    assert lin.item() < 1e-13 and pars.item() < 1e-12


GC_CASES = [
    dict(kind="gc", n=[8, 8, 8], np=[2, 2], gc_below=[0, 0, 0], gc_above=[4, 0, 0]),        # the reference test's halo
    dict(kind="gc", n=[8, 8, 8], np=[2, 2], gc_below=[2, 2, 0], gc_above=[3, 3, 0]),
    dict(kind="gc", n=[9, 7, 5], np=[2, 2], gc_below=[1, 2, 1], gc_above=[2, 1, 2]),        # ragged, halo in dim 2
    dict(kind="gc", n=[8, 6, 4], np=[4, 1], gc_below=[3, 0, 0], gc_above=[5, 0, 0]),        # halo wider than a block
    dict(kind="gc", n=[8, 8, 8], np=[2, 2], gc_below=[1, 1, 0], gc_above=[1, 1, 0], complex=False),
    dict(kind="gc", n=[8, 8, 8], np=[1, 1], gc_below=[2, 0, 1], gc_above=[1, 3, 0]),
]


@pytest.mark.parametrize("case", GC_CASES, ids=lambda c: "gc-%s-np%s-b%s-a%s" % (
    "x".join(map(str, c["n"])), "x".join(map(str, c["np"])), "".join(map(str, c["gc_below"])), "".join(map(str, c["gc_above"]))))
def test_ghost_cells_match_oracle(built_lib, case):
    """pfft_exchange / pfft_reduce against the oracle's net-effect definition."""
    import gpu_worker
    P = int(np.prod(case["np"]))
    with tempfile.TemporaryDirectory() as td:
        json.dump(case, open(os.path.join(td, "case.json"), "w"))
        cmd = [os.path.join(ROOT, "pfft_b200", "bin", "pfftrun"), "-np", str(P), "-timeout", "120",
               sys.executable, os.path.join(ROOT, "tests", "gpu_worker.py"), os.path.join(td, "case.json"), td]
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=200)
        assert p.returncode == 0, (p.stdout[-2000:], p.stderr[-2000:])
        res = []
        for rk in range(P):
            meta = json.load(open(os.path.join(td, "rank%d.json" % rk)))
            assert meta.get("error", "") == "", meta["error"]
            meta["out"] = np.load(os.path.join(td, "rank%d.npy" % rk))
            res.append(meta)
    n, gb, ga = case["n"], case["gc_below"], case["gc_above"]
    cplx = case.get("complex", True)
    rng = np.random.default_rng(5)
    xg = rng.standard_normal(n) + (1j * rng.standard_normal(n) if cplx else 0)
    blocks = []
    for r in res:
        tot = int(np.prod(r["ngc"]))
        ex = r["out"][:tot].reshape(r["ngc"])
        want = po.gc_exchange_block(xg, n, r["local_n"], r["local_start"], gb, ga)
        assert np.array_equal(ex, want), "exchange must be an exact copy"
        blocks.append((ex, r["local_n"], r["local_start"], gb, ga))
    summed = po.gc_reduce_global(n, blocks)
    for r in res:
        tot = int(np.prod(r["ngc"]))
        cnt = int(np.prod(r["local_n"]))
        red = r["out"][tot:tot + cnt].reshape(r["local_n"])
        sl = tuple(slice(r["local_start"][t], r["local_start"][t] + r["local_n"][t]) for t in range(3))
        assert np.allclose(red, summed[sl], rtol=1e-14, atol=1e-14)
        assert np.all(r["out"][tot + cnt:2 * tot] == 0), "tail must be zeroed after reduce"


# ---- BASELINE.json configs 3-5 at their full sizes (8 ranks; they share the GPU on a 1-GPU box)
PAD_C2R = po.PADDED_R2C   # PFFT_PADDED_C2R is the same bit (reference api/pfft.h:540-541)
BASELINE_CONFIGS = [
    dict(tag="config3-r2c-fp32-1024-padded-inplace", kind="r2c", n=[1024, 1024, 1024], np=[2, 4], precision="single",
         flags_forward=T_OUT | PAD, flags_backward=T_IN | PAD_C2R, inplace=True, tol=1500 * 2e-5),
    dict(tag="config4-c2c-128^4-on-2x2x2", kind="c2c", n=[128, 128, 128, 128], np=[2, 2, 2],
         flags_forward=T_OUT, flags_backward=T_IN, tol=1e-12),
    dict(tag="config5-ousam-r2c-512-to-768-gcell", kind="r2c", n=[768, 768, 768], ni=[512, 512, 512], np=[2, 4],
         flags_forward=T_OUT, flags_backward=T_IN, gc=dict(below=[2, 2, 0], above=[3, 3, 0]), tol=1e-12),
]


@pytest.mark.parametrize("cfg", BASELINE_CONFIGS, ids=lambda c: c["tag"])
def test_baseline_configs_round_trip_at_full_size(built_lib, cfg):
    P = int(np.prod(cfg["np"]))
    with tempfile.TemporaryDirectory() as td:
        json.dump(cfg, open(os.path.join(td, "cfg.json"), "w"))
        cmd = [os.path.join(ROOT, "pfft_b200", "bin", "pfftrun"), "-np", str(P), "-timeout", "500",
               sys.executable, os.path.join(ROOT, "tests", "baseline_worker.py"), os.path.join(td, "cfg.json"), td]
        p = subprocess.run(cmd, capture_output=True, text=True, timeout=560)
        assert p.returncode == 0, (p.returncode, p.stdout[-2000:], p.stderr[-2000:])
        res = [json.load(open(os.path.join(td, "rank%d.json" % rk))) for rk in range(P)]
    for r in res:
        assert r["error"] == "", r["error"]
    print("\n%s: kernels %s / %s, stage ms (rank 0, ranks share the GPU) forward %s backward %s" % (
        cfg["tag"], res[0]["kernels_forward"], res[0]["kernels_backward"],
        [round(v, 2) for v in res[0]["stage_ms_forward"]], [round(v, 2) for v in res[0]["stage_ms_backward"]]))
    # the reference's acceptance rule: pfft_check_output_* (a maximum over all ranks)
    assert max(r["maxerror"] for r in res) < cfg["tol"], [r["maxerror"] for r in res]
    # decomposition of the last rank as captured from the reference's own code
    if cfg["tag"].startswith("config3"):
        assert (res[7]["local_ni"], res[7]["local_i_start"], res[7]["local_no"]) == ([512, 256, 1026], [512, 768, 0], [1024, 512, 126])
    if cfg["tag"].startswith("config4"):
        assert (res[7]["local_ni"], res[7]["local_no"]) == ([64, 64, 64, 128], [128, 64, 64, 64])
    if cfg.get("gc"):
        assert (res[7]["local_ni"], res[7]["local_no"]) == ([256, 128, 512], [768, 384, 94])
        g7 = res[7]["gc"]
        assert (g7["ngc"], g7["gc_start"], g7["mem"]) == ([261, 133, 512], [254, 382, 0], 17773056)
        for r in res:
            g = r["gc"]
            assert g["error"] == "" and g["exchange_maxerr"] == 0.0 and g["tail_zero"]
        # reduce is the adjoint of exchange: what all ghost-cell blocks held is what the owners hold afterwards
        ex, red = sum(r["gc"]["exchanged_sum"] for r in res), sum(r["gc"]["reduced_sum"] for r in res)
        assert abs(ex - red) <= 1e-12 * abs(ex), (ex, red)


def test_execute_async_is_graph_capturable(world):
    """A stream-ordered execute (pfftb200_execute_async on device arrays) contains nothing but kernel launches with
    static parameters: captured once into a CUDA graph on the plan's stream and replayed on new data, it must
    reproduce the transform.  (Launch-bound small transforms are what graphs are for: 64^3 is three ~10 us kernels.)"""
    import torch
    pf = world
    n = [64, 64, 64]
    N = int(np.prod(n))
    comm = pf.create_procmesh([1, 1])
    side = torch.cuda.Stream()
    pf.set_stream(side.cuda_stream)
    try:
        x = torch.zeros(N, 2, dtype=torch.float64, device="cuda")
        X = torch.zeros_like(x)
        plan = pf.plan_dft(n, x, X, comm, pf.FORWARD, T_OUT, dtype=np.complex128)
        assert plan is not None, pf.last_error()
        plan.enable_stage_timing(False)
        g = torch.Generator(device="cuda").manual_seed(3)
        with torch.cuda.stream(side):
            x.copy_(torch.rand(N, 2, dtype=torch.float64, device="cuda", generator=g))
            plan.execute_async(x, X)                      # warm-up: buffers and kernel attributes are set up here
        side.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph, stream=side):
            plan.execute_async(x, X)
        for seed in (5, 6):
            fresh = torch.rand(N, 2, dtype=torch.float64, device="cuda", generator=g.manual_seed(seed))
            x.copy_(fresh)
            X.zero_()
            torch.cuda.synchronize()
            graph.replay()
            torch.cuda.synchronize()
            want = torch.fft.fftn(torch.view_as_complex(fresh).reshape(n)).permute(1, 2, 0).contiguous()
            got = torch.view_as_complex(X).reshape(n[1], n[2], n[0])
            err = ((got - want).abs().max() / want.abs().max()).item()
            assert err < 1e-12, err
        plan.destroy()
    finally:
        pf.set_stream(0)
        comm.free()

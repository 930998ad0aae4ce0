"""Oracle-checked parity on REAL multi-GPU hardware (-m gpu): one rank per GPU.

Every test here needs at least P visible GPUs and is SKIPPED (not passed) on a smaller box; ranks are
placed with rank -> GPU rank % device_count by pfftrun / LOCAL_RANK by torchrun, and each test asserts that
the ranks really sat on P distinct devices.  Both transports (peer-mapped stores, grouped NCCL all-to-all)
are exercised.  Semantics follow the reference's tests/simple_check_c2c_transposed.c (forward with
PFFT_TRANSPOSED_OUT, values compared in the transposed layout), with the numpy oracle on the gathered
array for sizes it finishes in seconds and, at BASELINE sizes (512^3 config 2, the headline 1024^3 on 2x4),
bench.py's spot check: K = 16 output coefficients against a direct fp64 summation of the hashed global input.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

import pfft_oracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu

T_IN, T_OUT, PAD = po.TRANSPOSED_IN, po.TRANSPOSED_OUT, po.PADDED_R2C


def ngpus():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=30).stdout
        return len([l for l in out.splitlines() if l.startswith("GPU ")])
    except Exception:
        return 0


NG = ngpus()


def need(P):
    if NG < P:
        pytest.skip("needs %d GPUs, this box has %d" % (P, NG))


FULL_GATHER = [
    dict(kind="c2c", n=[128, 128, 128], np=[2, 1], flags=T_OUT),
    dict(kind="c2c", n=[128, 128, 128], np=[2, 1], flags=T_IN, sign=+1),
    dict(kind="c2c", n=[256, 256, 256], np=[2, 1], flags=T_OUT),
    dict(kind="c2c", n=[29, 27, 31], np=[2, 1]),
    dict(kind="r2c", n=[64, 64, 128], np=[2, 1], flags=T_OUT | PAD, precision="single"),
    dict(kind="c2c", n=[128, 128, 128], np=[2, 2], flags=T_OUT),
    dict(kind="c2c", n=[256, 256, 256], np=[2, 2], flags=T_OUT),
    dict(kind="c2c", n=[29, 27, 31], np=[2, 2]),                       # BASELINE configs[0]
    dict(kind="c2r", n=[64, 64, 128], np=[2, 2], flags=T_IN, sign=+1),
    dict(kind="c2c", n=[128, 128, 128], np=[2, 4], flags=T_OUT),
    dict(kind="c2c", n=[128, 128, 128], np=[2, 4], flags=T_IN, sign=+1),
    dict(kind="c2c", n=[256, 256, 256], np=[2, 4], flags=T_OUT),
    dict(kind="c2c", n=[29, 27, 31], np=[2, 4], flags=T_OUT),
    dict(kind="c2c", n=[32, 32, 32, 32], np=[2, 2, 2], flags=T_OUT),
    dict(kind="r2c", n=[96, 96, 96], ni=[64, 64, 64], no=[96, 96, 96], np=[2, 4], flags=T_OUT),
]


def _id(c):
    return "%s-%s-np%s-f%d-%s" % (c["kind"], "x".join(map(str, c["n"])), "x".join(map(str, c["np"])), c.get("flags", 0),
                                  c.get("precision", "double")[0])


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
@pytest.mark.parametrize("case", FULL_GATHER, ids=_id)
def test_one_rank_per_gpu_matches_oracle(built_lib, case, transport):
    P = int(np.prod(case["np"]))
    need(P)
    import test_gpu_parity as tp
    (err, l2), results = tp.run_multi(case, env_extra={"PFFT_B200_TRANSPORT": transport})
    assert sorted(r["device"] for r in results) == list(range(P)), "ranks must sit on distinct GPUs"
    assert all(r["transport"] == transport for r in results)
    prec = case.get("precision", "double")
    assert err < tp.TOL[prec], err
    assert l2 < tp.TOL_L2[prec], l2


def run_bench(P, extra, timeout=900):
    port = 29500 + (os.getpid() % 400)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(P), "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "bench.py"), "--gpus", str(P), "--steps", "2", "--warmup", "1",
           "--no-cpu", "--no-e2e"] + extra
    if P == 1:
        cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--gpus", "1", "--steps", "2", "--warmup", "1", "--no-cpu", "--no-e2e"] + extra
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    assert p.returncode == 0, (p.stdout[-3000:], p.stderr[-3000:])
    lines = [l for l in p.stdout.splitlines() if l.startswith("{")]
    assert lines, p.stdout[-2000:]
    return json.loads(lines[-1])


SPOT = [
    # (ranks, bench config, mesh, transport): config 2 = 512^3 on its pencil meshes, config 1 = the headline 1024^3
    (2, 2, "2x1", "p2p"), (2, 2, "2x1", "nccl"), (2, 1, "2x1", "p2p"),
    (4, 2, "2x2", "p2p"), (4, 2, "2x2", "nccl"), (4, 1, "2x2", "p2p"),
    (8, 2, "2x4", "p2p"), (8, 2, "2x4", "nccl"), (8, 1, "2x4", "p2p"), (8, 1, "2x4", "nccl"),
    (8, 3, "2x4", "p2p"), (8, 4, "2x2x2", "p2p"), (8, 5, "2x4", "p2p"),
]


@pytest.mark.parametrize("P,config,mesh,transport", SPOT, ids=lambda v: str(v))
def test_baseline_size_forward_values_by_direct_summation(built_lib, P, config, mesh, transport):
    """The benchmark's own transform (same plans, same mesh, same transport as the timed path): 16 forward
    coefficients against a direct summation, and the forward+backward round trip, relative L2."""
    need(P)
    line = run_bench(P, ["--config", str(config), "--mesh", mesh, "--transport", transport])
    assert line["n_gpus"] == P and line["transport"] == transport
    f32 = line["dtype"] == "f32"
    sc = line["spot_check"]
    assert sc["coefficients"] >= 16
    assert sc["max_err_over_rms"] < (2e-5 if f32 else 1e-12), sc
    assert line["roundtrip_rel_l2_err"] < (1e-5 if f32 else 1e-12), line["roundtrip_rel_l2_err"]


def test_spot_check_on_one_gpu_headline(built_lib):
    """Same check on a single GPU (always runs): the headline 1024^3 c2c fp64."""
    line = run_bench(1, ["--config", "1"])
    assert line["spot_check"]["max_err_over_rms"] < 1e-12, line["spot_check"]
    assert line["roundtrip_rel_l2_err"] < 1e-12

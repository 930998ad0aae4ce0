"""bench.py's input generator and spot-check arithmetic against the oracle (CPU, no GPU needed):
the counter-hash block generated with torch must equal oracle/pfft_oracle.py: synthetic_complex bit for
bit, and the direct summation must reproduce numpy's transform of the same array."""
import os
import sys

import numpy as np
import pytest
import torch

import pfft_oracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


@pytest.mark.parametrize("n,ln,ls", [([8, 6, 10], [8, 6, 10], [0, 0, 0]), ([16, 12, 10], [5, 7, 10], [11, 3, 0]),
                                     ([6, 5, 4, 3], [3, 2, 4, 3], [3, 1, 0, 0])])
def test_hashed_block_equals_oracle(n, ln, ls):
    want = po.synthetic_complex(n, ln, ls)
    got = bench.synthetic_block(torch, n, ln, ls, False, torch.float64, torch.device("cpu"))
    got = torch.view_as_complex(got).numpy()
    assert np.array_equal(got, want)
    gr = bench.synthetic_block(torch, n, ln, ls, True, torch.float64, torch.device("cpu"), row_pitch=ln[-1] + 2).numpy()
    assert np.array_equal(gr[..., :ln[-1]], want.real) and np.all(gr[..., ln[-1]:] == 0)


def test_direct_sums_match_numpy_fft():
    n = [8, 6, 10]
    x = po.synthetic_complex(n, n, [0, 0, 0])
    X = np.fft.fftn(x)
    ks = [[0, 0, 0], [7, 5, 9], [3, 2, 1], [1, 0, 4]]
    # two "ranks" holding halves of dimension 0
    tot = 0
    for start, cnt in ((0, 5), (5, 3)):
        blk = torch.view_as_real(torch.from_numpy(np.ascontiguousarray(x[start:start + cnt])))
        tot = tot + bench.direct_partial_sums(torch, blk, n, [start, 0, 0], ks, False).numpy()
    want = np.array([X[tuple(k)] for k in ks])
    assert np.abs(tot - want).max() < 1e-12
    # real input, pruned: ni = (8, 6, 10) zero-padded to n = (12, 9, 16)
    nn = [12, 9, 16]
    xr = x.real
    Xr = np.fft.rfftn(xr, s=nn, axes=[0, 1, 2])
    ks = [[0, 0, 0], [11, 8, 8], [5, 4, 3]]
    got = bench.direct_partial_sums(torch, torch.from_numpy(np.ascontiguousarray(xr)), nn, [0, 0, 0], ks, True).numpy()
    assert np.abs(got - np.array([Xr[tuple(k)] for k in ks])).max() < 1e-12


def test_reference_arm_is_independent_of_torchrun_thread_env(tmp_path):
    """torchrun exports OMP_NUM_THREADS=1 for nproc > 1; the CPU arm must pin its own pool size."""
    import json
    import subprocess
    outs = []
    for env_threads in (None, "1"):
        env = dict(os.environ)
        env.pop("OMP_NUM_THREADS", None)
        if env_threads:
            env["OMP_NUM_THREADS"] = env_threads
        p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "2", "--size", "128",
                            "--steps", "3", "--warmup", "1"], capture_output=True, text=True, env=env, timeout=300)
        assert p.returncode == 0, p.stderr[-2000:]
        outs.append(json.loads(p.stdout.strip().splitlines()[-1]))
    for o in outs:
        assert o["impl"] == "reference" and o["steps"] == 3 and o["config"]["cpu_sample_n"] == [128, 128, 128]
        assert o["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    if len(os.sched_getaffinity(0)) >= 4:
        a, b = outs[0]["value"], outs[1]["value"]
        assert min(a, b) / max(a, b) > 0.4, (a, b)     # (an 8-core pool throttled to one thread is 5x slower)

"""CPU tests of the drop-in boundary: the shared library loads without a GPU and exports every
symbol the public headers declare; the minimpi launcher runs a 4-rank C program; a 2-rank
torch.distributed (gloo) job drives the integer API through the Python host layer."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
INC = os.path.join(ROOT, "include")


def declared_symbols():
    names = set()
    pf = open(os.path.join(INC, "pfft.h")).read()
    body = pf[pf.index("#define PFFT_B200_DECLARE_API"):pf.index("PFFT_B200_DECLARE_API(PFFT_MANGLE_DOUBLE")]
    for m in re.finditer(r"\bP\((\w+)\)\s*\(", body):
        for prefix in ("pfft_", "pfftf_"):
            names.add(prefix + m.group(1))
    for hdr, pat in (("mpi.h", r"\b(MPI_\w+|minimpi_\w+)\s*\("), ("pfft_b200.h", r"\b(pfftb200_\w+)\s*\("),
                     ("fftw3.h", r"\b(fftwf?_(?:malloc|alloc_real|alloc_complex|free|plan_dft_3d|execute|destroy_plan))\s*\("),
                     ("fftw3-mpi.h", r"\b(fftw_mpi_\w+)\s*\(")):
        text = open(os.path.join(INC, hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        text = "\n".join(l for l in text.splitlines() if not l.lstrip().startswith("#"))   # macros are not symbols
        for m in re.finditer(pat, text):
            names.add(m.group(1))
    names.discard("minimpi_comm_s")
    return sorted(names)


def test_library_exports_every_declared_symbol(built_lib):
    names = declared_symbols()
    assert len(names) > 300
    missing = [n for n in names if not hasattr(built_lib, n)]
    assert not missing, missing


def test_version_and_describe_without_gpu(built_lib):
    import pfft_b200 as pf
    assert "sm_100a" in pf.version()
    s = pf.describe_schedule("c2c", [1024, 1024, 1024], [2, 4], 7, pf.TRANSPOSED_OUT)
    assert s["local_ni"] == [512, 256, 1024] and s["local_no"] == [1024, 512, 256]   # SURVEY.md 8 (pid 7)
    assert s["local_i_start"] == [512, 768, 0] and s["local_o_start"] == [0, 512, 768]


def test_minimpi_four_ranks_c_program(built_lib, tmp_path):
    exe = str(tmp_path / "mpi_smoke")
    subprocess.check_call(["gcc", "-std=gnu99", "-O1", "-I" + INC, os.path.join(ROOT, "tests", "c", "mpi_smoke.c"),
                           "-o", exe, "-L" + os.path.join(ROOT, "pfft_b200", "lib"), "-lpfft_b200",
                           "-Wl,-rpath," + os.path.join(ROOT, "pfft_b200", "lib")])
    p = subprocess.run([os.path.join(ROOT, "pfft_b200", "bin", "pfftrun"), "-np", "4", "-timeout", "60", exe],
                       capture_output=True, text=True, timeout=90)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "ranks=4 sum=10 max=4" in p.stdout and "ok=1" in p.stdout
    # golden TRANSPOSED_OUT blocks of 29x27x31 on 2x2
    assert "rank 3 coords (1,1) ni=[14,13,31] no=[29,13,15] os=[0,14,16]" in p.stdout


WORKER = r'''
import os, sys, json
sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, "oracle"))
import torch.distributed as dist
dist.init_process_group("gloo")
import pfft_b200 as pf, pfft_oracle as po
pf.init()
rank, size = dist.get_rank(), dist.get_world_size()
comm = pf.create_procmesh([2, 1])
assert comm.rank == rank and comm.size == size
out = {{}}
for kind, n, flags in (("c2c", [29, 27, 31], 0), ("c2c", [29, 27, 31], pf.TRANSPOSED_OUT),
                       ("r2c", [16, 12, 10], pf.TRANSPOSED_OUT | pf.PADDED_R2C), ("c2r", [16, 12, 10], pf.TRANSPOSED_IN)):
    alloc, lni, lis, lno, los = pf.local_size(kind, n, comm, flags)
    want = po.local_block(kind, n, n, [2, 1], rank, flags)
    assert (lni, lis, lno, los) == tuple(map(list, want)), (kind, n, flags)
    assert alloc >= 1
    sched = pf.describe_schedule(kind, n, [2, 1], rank, flags)
    assert sched["error"] == "" and len(sched["exchanges"]) >= 1
assert abs(comm.allreduce_max(float(rank)) - (size - 1)) < 1e-15
comm.barrier()
comm.free(); pf.finalize(); dist.destroy_process_group()
sys.stdout.write("worker%d-ok\n" % rank)
'''


def test_two_ranks_gloo_host_logic(built_lib, tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ)
    env.pop("PFFT_MPI_JOB", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=240, env=env)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    assert "worker0-ok" in p.stdout and "worker1-ok" in p.stdout

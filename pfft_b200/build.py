"""Build recipe for libpfft_b200.so (in-tree, sm_100a only).

    python -m pfft_b200.build            # or: __graft_entry__.build()

Everything is compiled by nvcc with `-gencode arch=compute_100a,code=sm_100a -lineinfo`
and linked into one shared library that exports the PFFT C API (include/pfft.h), the
minimpi subset (include/mpi.h) and the pfftb200_* extensions (include/pfft_b200.h).
Also builds the `pfftrun` launcher.  NCCL is loaded at run time with dlopen.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
BUILD = os.path.join(ROOT, "build", "obj")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")

SOURCES = ["minimpi.cpp", "decomp.cpp", "planner.cpp", "describe.cpp", "fft_tables.cpp",
           "fft_generic.cu", "fft_mixed.cu", "fft_pow2.cu", "fft_reg.cu",
           "fft_reg_k0_f64.cu", "fft_reg_k1_f64.cu", "fft_reg_k2_f64.cu", "fft_reg_k0_f32.cu", "fft_reg_k1_f32.cu", "fft_reg_k2_f32.cu", "plan.cu", "transports.cu", "gcell.cu", "api.cu"]
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
CFLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-Wall,-Wno-unused-function",
          "-I" + os.path.join(ROOT, "include"), "-I" + CSRC, "-I/usr/include"]


def _newer(src, dst, extra=()):
    if not os.path.exists(dst):
        return True
    t = os.path.getmtime(dst)
    return any(os.path.getmtime(p) > t for p in (src,) + tuple(extra))


def build(verbose=False, force=False):
    os.makedirs(LIBDIR, exist_ok=True)
    os.makedirs(BUILD, exist_ok=True)
    headers = tuple(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")) + tuple(
        os.path.join(ROOT, "include", f) for f in os.listdir(os.path.join(ROOT, "include")))
    objs = []
    procs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(BUILD, s + ".o")
        objs.append(obj)
        if force or _newer(src, obj, headers):
            cmd = [NVCC] + ARCH + CFLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-x", "cu", "-c", src, "-o", obj]
            procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for s, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write("---- %s\n%s\n" % (s, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    lib = os.path.join(LIBDIR, "libpfft_b200.so")
    if force or procs or not os.path.exists(lib):
        cmd = [NVCC] + ARCH + ["-shared", "-o", lib] + objs + ["-lcudart", "-ldl", "-lrt", "-lpthread"]
        subprocess.check_call(cmd)
    run = os.path.join(HERE, "bin", "pfftrun")
    src = os.path.join(HERE, "tools", "pfftrun.c")
    os.makedirs(os.path.dirname(run), exist_ok=True)
    if os.path.exists(src) and (force or _newer(src, run)):
        subprocess.check_call(["gcc", "-O2", "-Wall", "-o", run, src, "-lrt"])
    return lib


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))

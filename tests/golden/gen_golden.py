"""tests/golden/gen_golden.py -- regenerates the golden fixtures in this directory.

Runs ONLY in the build container (needs /root/reference): it calls the
reference's own integer code, compiled by `make -C oracle` into
oracle/_ref/libpfft_refint.so, and records its answers.  The fixtures are what
the GPU box (which has no /root/reference) tests against.

    make -C oracle && python tests/golden/gen_golden.py
"""
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "oracle"))
import refint  # noqa: E402

T_NONE, T_IN, T_OUT = 0, 1, 2
S_IN, S_OUT, PADDED = 4, 8, 1 << 11
KINDS = {"c2c": "dft", "r2c": "dft_r2c", "c2r": "dft_c2r", "r2r": "r2r"}


def main():
    # silence the reference's odd-size SHIFTED warnings
    devnull = os.open(os.devnull, os.O_WRONLY)
    os.dup2(devnull, 2)
    ref = refint.RefInt()
    rnd = random.Random(20261017)

    # ---- local_block golden vectors -------------------------------------------
    cases = []
    meshes = [[1], [2], [4], [8], [3], [1, 1], [2, 1], [1, 2], [2, 2], [2, 4], [4, 2], [3, 2], [8, 1],
              [2, 2, 2], [1, 2, 4], [2, 3, 2], [1, 1, 8]]
    sizes = [[29, 27, 31], [16, 16, 16], [5, 3, 2], [512, 512, 512], [1024, 1024, 1024], [768, 768, 768],
             [13, 14, 19, 17], [128, 128, 128, 128], [8, 8, 8, 8], [16, 9], [7, 1, 3]]
    for mesh in meshes:
        for n in sizes:
            d = len(n)
            if d < len(mesh) or (d == len(mesh) and d != 3):
                continue
            for kind in KINDS:
                for tf in (T_NONE, T_IN, T_OUT):
                    if (kind == "r2c" and tf == T_IN) or (kind == "c2r" and tf == T_OUT):
                        continue
                    for extra in (0, PADDED, S_IN | S_OUT):
                        if extra == PADDED and kind in ("c2c", "r2r"):
                            continue
                        if extra & S_IN and any(x % 2 for x in n):
                            continue
                        variants = [(n, n, None, None)]
                        if extra == 0 and max(n) <= 128:
                            ni = [max(2, x - rnd.randint(0, x // 2)) for x in n]
                            no = [max(2, x - rnd.randint(0, x // 2)) for x in n]
                            variants.append((ni, no, None, None))
                            r = min(len(mesh), d - 1) if not (d == 3 and len(mesh) == 3) else None
                            if r:
                                ib = [rnd.randint(1, max(1, n[t])) for t in range(r)]
                                ob = [rnd.randint(1, max(1, n[t])) for t in range(r)]
                                variants.append((n, n, ib, ob))
                        P = 1
                        for p in mesh:
                            P *= p
                        pids = range(P) if (P <= 4 and max(n) <= 32) else sorted({0, P // 2, P - 1})
                        for (ni, no, ib, ob) in variants:
                            for pid in pids:
                                out = ref.local_block(KINDS[kind], ni, no, mesh, pid, tf | extra, ib, ob)
                                cases.append([kind, ni, no, mesh, pid, tf | extra, ib, ob,
                                              out[0], out[1], out[2], out[3]])
    with open(os.path.join(HERE, "local_block.json"), "w") as f:
        json.dump({"source": "pfft_local_block_many_* of the reference (kernel/partrafo.c:99-199)",
                   "columns": ["kind", "ni", "no", "np", "pid", "flags", "iblock", "oblock",
                               "lni", "lis", "lno", "los"],
                   "cases": cases}, f, separators=(",", ":"))

    # ---- ghost-cell sizes -------------------------------------------------------
    gcs = []
    for _ in range(200):
        ln = [rnd.randint(1, 300) for _ in range(3)]
        ls = [rnd.randint(0, 700) for _ in range(3)]
        gb = [rnd.randint(0, 5) for _ in range(3)]
        ga = [rnd.randint(0, 5) for _ in range(3)]
        hm = rnd.choice([1, 1, 2, 3])
        mem, ngc, gstart = ref.local_size_gc(ln, ls, hm, gb, ga)
        gcs.append({"local_n": ln, "local_start": ls, "howmany": hm, "gc_below": gb, "gc_above": ga,
                    "mem": mem, "ngc": ngc, "gc_start": gstart})
    mem, ngc, gstart = ref.local_size_gc([256, 128, 512], [256, 384, 0], 1, [2, 2, 0], [3, 3, 0])
    gcs.append({"local_n": [256, 128, 512], "local_start": [256, 384, 0], "howmany": 1,
                "gc_below": [2, 2, 0], "gc_above": [3, 3, 0], "mem": mem, "ngc": ngc, "gc_start": gstart})
    with open(os.path.join(HERE, "local_size_gc.json"), "w") as f:
        json.dump({"source": "pfft_local_size_many_gc (gcell/gcells_plan.c:51-76)", "cases": gcs}, f,
                  separators=(",", ":"))

    # ---- init_input patterns (bit-exact doubles stored as hex) ---------------------
    pats = []
    specs = [("complex", [29, 27, 31], [2, 2, 3], [0, 0, 0]), ("complex", [29, 27, 31], [3, 2, 4], [15, 14, 0]),
             ("complex", [29, 27, 31], [2, 3, 5], [27, 24, 26]), ("real", [4, 4, 3], [1, 1, 4], [0, 0, 0]),
             ("real", [16, 16, 16], [2, 3, 18], [8, 8, 0]), ("complex_hermitian", [16, 16, 16], [3, 2, 9], [8, 8, 0]),
             ("complex_hermitian", [29, 27, 31], [2, 2, 16], [0, 0, 0]),
             ("complex", [16, 16, 16], [2, 2, 16], [0, 0, -8]), ("complex", [13, 14, 19, 17], [2, 2, 2, 3], [7, 7, 10, 9]),
             ("complex", [8, 8, 8], [2, 2, 9], [6, 7, 0])]
    for kind, n, ln, ls in specs:
        buf = ref.init_input(kind, n, ln, ls)
        if kind == "real":
            vals = [float(x).hex() for x in buf]
        else:
            vals = [[float(x.real).hex(), float(x.imag).hex()] for x in buf]
        pats.append({"kind": kind, "n": n, "local_n": ln, "local_start": ls, "values": vals})
    with open(os.path.join(HERE, "init_input.json"), "w") as f:
        json.dump({"source": "pfft_init_input_* (api/api-basic.c:60-121)", "cases": pats}, f,
                  separators=(",", ":"))
    print(len(cases), "local_block cases;", len(gcs), "gc cases;", len(pats), "patterns")


if __name__ == "__main__":
    main()

"""Time BASELINE.json configs 2-5 at full size on the GPUs of this box (one rank per GPU under pfftrun):
   python profiles/time_baseline_configs.py [nranks]   -> table on stdout.
Uses tests/baseline_worker.py (round trip checked first, then timed forward+backward pairs)."""
import json
import math
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T_IN, T_OUT, PAD = 1 << 0, 1 << 1, 1 << 11      # PFFT_TRANSPOSED_IN / _OUT, PFFT_PADDED_R2C (include/pfft.h)
CONFIGS = [
    dict(tag="config2 c2c fp64 512^3 2x4", kind="c2c", n=[512] * 3, np=[2, 4], flags_forward=T_OUT, flags_backward=T_IN),
    dict(tag="config3 r2c/c2r fp32 1024^3 padded in place 2x4", kind="r2c", n=[1024] * 3, np=[2, 4], precision="single",
         flags_forward=T_OUT | PAD, flags_backward=T_IN | PAD, inplace=True),
    dict(tag="config4 c2c fp64 128^4 on 2x2x2", kind="c2c", n=[128] * 4, np=[2, 2, 2], flags_forward=T_OUT, flags_backward=T_IN),
    dict(tag="config5 r2c/c2r fp64 512^3 -> 768^3 (ousam) 2x4", kind="r2c", n=[768] * 3, ni=[512] * 3, np=[2, 4],
         flags_forward=T_OUT, flags_backward=T_IN),
]


def main():
    single = len(sys.argv) > 1 and sys.argv[1] == "1"      # one rank on a 1x1 mesh: kernel throughput without exchanges
    only = sys.argv[2] if len(sys.argv) > 2 else ""          # substring filter on the tag
    for cfg in CONFIGS:
        if only not in cfg["tag"]:
            continue
        cfg = dict(cfg, time_pairs=5)
        if single:
            cfg["np"] = [1] * len(cfg["np"])
            cfg["tag"] = cfg["tag"].rsplit(" ", 1)[0] + " 1 GPU"

        P = 1
        for v in cfg["np"]:
            P *= v
        with tempfile.TemporaryDirectory() as td:
            json.dump(cfg, open(os.path.join(td, "cfg.json"), "w"))
            cmd = [os.path.join(ROOT, "pfft_b200", "bin", "pfftrun"), "-np", str(P), "-timeout", "300", sys.executable,
                   os.path.join(ROOT, "tests", "baseline_worker.py"), os.path.join(td, "cfg.json"), td]
            p = subprocess.run(cmd, capture_output=True, text=True, timeout=360)
            if p.returncode != 0:
                print("%s: FAILED rc=%d %s" % (cfg["tag"], p.returncode, p.stderr[-500:]))
                continue
            r = json.load(open(os.path.join(td, "rank0.json")))
        N = 1
        for v in cfg["n"]:
            N *= v
        flop = (5.0 if cfg["kind"] == "c2c" else 2.5) * N * math.log2(N) * 2      # forward + backward
        print("%-52s maxerror %.1e  %8.3f ms/pair  %8.1f GFlop/s  kernels %s  stage ms fwd %s bwd %s" % (
            cfg["tag"], r["maxerror"], r["ms_per_pair"], flop / (r["ms_per_pair"] * 1e-3) / 1e9, r["kernels_forward"],
            [round(v, 2) for v in r["stage_ms_forward"]], [round(v, 2) for v in r["stage_ms_backward"]]))
        sys.stdout.flush()


if __name__ == "__main__":
    main()

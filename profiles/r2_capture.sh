#!/bin/bash
# Round-2 profile captures on ONE B200 (run under gpurun).  The .ncu-rep files stay on the box (too large to
# bring back); what comes back under gpurun_out/ are their exports: selected raw metrics per launch
# (profiles/ncu_metrics.py) and the source-level stall summary (profiles/ncu_source_stalls.py).
#   1. launch list of the default bench command (gpu__time_duration.sum per launch)
#   2. ncu --set full of every stage kernel family on the BASELINE configs:
#      config 1 (stage_blk_kernel), 3 (stage_reg_kernel r2c/c2r fp32, fused_pair_kernel), 5 (stage_reg_kernel 768 / 384,
#      gc_gather / gc_reduce), and odd sizes (stage_mixed_kernel: 232 x 216 x 248 = 8 x the reference test's 29 x 27 x 31)
mkdir -p gpurun_out
R=/tmp/ncu_reps; mkdir -p $R
N="ncu --clock-control none"
timeout 300 $N --metrics gpu__time_duration.sum -k "regex:stage_|fused_pair|gc_|xch_" -c 400 --csv --log-file gpurun_out/p2_launches_bench_n1.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-spot > gpurun_out/p2_launches.log 2>&1   # (library kernels only: the input generation alone is thousands of torch launches)
python profiles/summarize_launches.py gpurun_out/p2_launches_bench_n1.csv > gpurun_out/p2_launches_bench_n1.txt 2>&1
cap() {  # name, kernel regex, count, command...
  name=$1; rx=$2; cnt=$3; shift 3
  timeout 240 $N --set full --import-source on -k "regex:$rx" -c $cnt -o $R/$name -f "$@" > gpurun_out/p2_${name}_run.log 2>&1
  timeout 120 python profiles/ncu_metrics.py $R/$name.ncu-rep > gpurun_out/p2_ncu_full_$name.txt 2>&1
  timeout 120 python profiles/ncu_source_stalls.py $R/$name.ncu-rep 30 > gpurun_out/p2_ncu_source_$name.txt 2>&1
}
cap blk_c1 "stage_blk" 3 python bench.py --steps 1 --warmup 0 --no-cpu --no-e2e --no-spot
cap c3 "stage_reg|fused_pair" 6 python bench.py --config 3 --steps 1 --warmup 0 --no-cpu --no-e2e --no-spot
cap c5 "stage_reg" 6 python bench.py --config 5 --steps 1 --warmup 0 --no-cpu --no-e2e --no-spot
cap gc "gc_" 4 python bench.py --config 5 --steps 1 --warmup 0 --no-cpu --no-e2e --no-spot
cap mixed "stage_mixed|stage_generic" 3 python -c "
import sys; sys.path[:0]=['tests','oracle','.']
import pfft_b200 as pf, gpu_worker
pf.init()
case=dict(kind='c2c', n=[29*8,27*8,31*8], np=[1,1])
comm=pf.create_procmesh(case['np']); r=gpu_worker.run_case(case, comm); print(r['kernels'], r['error'])
"
du -sh gpurun_out

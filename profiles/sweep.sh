#!/bin/bash
# usage: profiles/sweep.sh "ENV1=a ENV2=b" "ENV1=c" ...   -- one short bench run per environment setting
mkdir -p gpurun_out
for v in "$@"; do
  echo "== $v" | tee -a gpurun_out/sweep.log
  env $v timeout 300 python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu 2>&1 | python profiles/print_bench.py | tee -a gpurun_out/sweep.log
done

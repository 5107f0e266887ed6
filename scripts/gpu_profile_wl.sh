#!/bin/bash
# full ncu capture of the dominant kernel of each workload in $WLS (default: c3 c4 c5); plain bench first
mkdir -p gpurun_out
for w in ${WLS:-c3 c4 c5}; do
  python bench.py --steps ${STEPS:-20} --warmup 3 --no-cpu-baseline --workload $w > gpurun_out/bench_$w.log 2> gpurun_out/bench_$w.err
  ncu --set full --clock-control none --import-source on -k regex:tmm_ -s 3 -c 1 -f -o gpurun_out/prof_$w \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --workload $w > gpurun_out/ncu_$w.log 2>&1
  tail -n 1 gpurun_out/ncu_$w.log
done
python scripts/show_bench.py

#!/bin/bash
# quick iteration: GPU tests + short bench (no CPU baseline) [+ one full ncu capture when PROFILE=1]
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
A="python bench.py --steps ${STEPS:-50} --warmup 3 --no-cpu-baseline --no-configs --no-sustained --no-ablation ${BENCH_ARGS:-}"
$A > gpurun_out/bench_quick.log 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python scripts/show_bench.py gpurun_out/bench_quick.log 2>/dev/null || tail -c 1500 gpurun_out/bench_quick.log
if [ "${PROFILE:-0}" = "1" ]; then
  B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-configs --no-sustained --no-ablation ${BENCH_ARGS:-}"
  ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-tmm_} -s 3 -c 1 -f -o gpurun_out/${OUT:-prof} $B > gpurun_out/ncu_quick.log 2>&1
  tail -n 2 gpurun_out/ncu_quick.log
fi

#!/bin/bash
# Round-2 evidence pass on one GPU: GPU tests (all of them, no -x), then for the bench command the ncu launch list and
# one full capture each of the C2 cavity kernel, the C2 layer loop (--no-layer-cse, K1 untyped) and the C4 scan kernel.
mkdir -p gpurun_out
if [ "${TESTS:-1}" = "1" ]; then
  ( timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  tail -5 gpurun_out/pytest_gpu.log
fi
Q="--steps 3 --warmup 3 --no-cpu-baseline --no-configs --no-sustained --no-ablation"
CMD="python bench.py $Q"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
cap() {  # name, kernel regex, bench args
  python bench.py $Q $3 > gpurun_out/plain_$1.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$2 -s 3 -c 1 -f -o gpurun_out/prof_$1 python bench.py $Q $3 > gpurun_out/ncu_$1.log 2>&1
  tail -n 1 gpurun_out/ncu_$1.log
}
for w in ${CAPS:-c2 nocse c4}; do
  case $w in
    c2) cap c2 tmm_cavity "" ;;
    nocse) cap nocse tmm_chain "--no-layer-cse" ;;
    c4) cap c4 tmm_chain "--workload c4" ;;
    c3) cap c3 tmm_sweep "--workload c3" ;;
    c5) cap c5 tmm_sweep "--workload c5" ;;
  esac
done
ls -la gpurun_out | head -40

#!/bin/bash
# Build-and-bench loop on the GPU box: one bench line per compile-time variant (PST_NVCC_EXTRA flags).
#   VARIANTS="-DPST_CAV_MINB2=6|-DPST_CAV_MINB2=7" BENCH_ARGS="--workload c4" bash scripts/gpu_variants.sh
mkdir -p gpurun_out
IFS='|' read -ra V <<< "${VARIANTS:-}"
for v in "${V[@]}"; do
  PST_NVCC_EXTRA="$v" python -m pistachio_b200.build --force --verbose 2>&1 | grep -A2 "${KGREP:-tmm_cavity_kernelILi2E}" | grep -E "spill|registers" | head -4
  python bench.py --steps ${STEPS:-100} --warmup 3 --no-cpu-baseline --no-configs --no-sustained --no-ablation ${BENCH_ARGS:-} > gpurun_out/var.log 2> gpurun_out/var.err
  echo "variant [$v]: $(python -c "
import json; d=json.loads(open('gpurun_out/var.log').read().strip().splitlines()[-1]); print('value %.4g ms %.4f' % (d['value'], d['ms_per_step']))" 2>&1 | tail -1)"
done
python -m pistachio_b200.build --force > /dev/null

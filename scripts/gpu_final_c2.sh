#!/bin/bash
# Round-end evidence for the bench workload only: smoke, GPU tests, default bench (with CPU baseline), reference arm,
# launch list + one full capture of the typed chain kernel.
mkdir -p gpurun_out
STEPS=200 bash scripts/gpu_check.sh > gpurun_out/check.out 2>&1
tail -2 gpurun_out/smoke.log; tail -3 gpurun_out/pytest_gpu.log
( timeout 300 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
bash scripts/gpu_profile.sh > gpurun_out/profile.out 2>&1
python scripts/show_bench.py gpurun_out/bench.log; tail -c 400 gpurun_out/bench_ref.log; ls -la gpurun_out | tail -20

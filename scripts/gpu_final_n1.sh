#!/bin/bash
# Round-end evidence on one GPU: smoke, GPU tests, default bench (with CPU baseline), reference arm, launch list + full
# captures of the dominant kernel of every workload, full-size C3/C5, fp32 mode.
mkdir -p gpurun_out
STEPS=200 bash scripts/gpu_check.sh > gpurun_out/check.out 2>&1
tail -2 gpurun_out/smoke.log; tail -3 gpurun_out/pytest_gpu.log
( timeout 300 python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/bench_ref.log 2> gpurun_out/bench_ref.err
bash scripts/gpu_profile.sh > gpurun_out/profile.out 2>&1
WLS="c3 c4 c5 c1" STEPS=50 bash scripts/gpu_profile_wl.sh > gpurun_out/profile_wl.out 2>&1
python bench.py --steps 50 --no-cpu-baseline --no-layer-cse > gpurun_out/bench_c2nocse.log 2>&1
python bench.py --steps 50 --no-cpu-baseline --fp32 > gpurun_out/bench_c2f32.log 2>&1
python scripts/full_configs.py > gpurun_out/full_n1.jsonl 2> gpurun_out/full_n1.err
python scripts/show_bench.py; cat gpurun_out/full_n1.jsonl | cut -c1-600

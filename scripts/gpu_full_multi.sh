#!/bin/bash
# full-size C5 (and C3) at N GPUs, as torchrun launches it
mkdir -p gpurun_out
N=${NGPU:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/full_configs.py --configs ${CONFIGS:-c5} > gpurun_out/full_n$N.jsonl 2> gpurun_out/full_n$N.err
echo "rc=$?"; cat gpurun_out/full_n$N.jsonl; grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/full_n$N.err | tail -15

#!/bin/bash
# multi-GPU bench as the driver launches it (N from $NGPU)
mkdir -p gpurun_out
N=${NGPU:-2}
nvidia-smi -L > gpurun_out/gpus.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 30 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
echo "rc=$?" >> gpurun_out/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 2 --warmup 1 --impl reference > gpurun_out/bench_ref_n$N.log 2> gpurun_out/bench_ref_n$N.err
echo "rc=$?" >> gpurun_out/bench_ref_n$N.err
tail -c 1500 gpurun_out/bench_n$N.log; tail -5 gpurun_out/bench_n$N.err; tail -c 600 gpurun_out/bench_ref_n$N.log

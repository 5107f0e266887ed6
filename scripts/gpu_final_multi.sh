#!/bin/bash
# Round-end multi-GPU evidence at N = $NGPU: the bench as the driver launches it (ours + reference arm) and the
# full-size C5 strong-scaling run with NCCL and fused assembly.
mkdir -p gpurun_out
N=${NGPU:-2}
nvidia-smi -L > gpurun_out/gpus_n$N.txt
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 100 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err
echo "bench rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 2 --warmup 1 --impl reference > gpurun_out/bench_ref_n$N.log 2> gpurun_out/bench_ref_n$N.err
echo "ref rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/full_configs.py --configs ${CONFIGS:-c5 c3} > gpurun_out/full_n$N.jsonl 2> gpurun_out/full_n$N.err
echo "full rc=$?"
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_n$N.log').read().strip().splitlines()[-1])
print('N=$N value %.4g ms %.4g e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value']), d.get('allgather'))
for l in open('gpurun_out/full_n$N.jsonl'):
    print(l[:900])
PY
grep -v "^W\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/bench_n$N.err | tail -3

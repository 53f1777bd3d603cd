#!/usr/bin/env bash
# Multi-GPU pass (gpurun --gpus N): the pure copy ceiling and the bench at 1 .. N ranks.
set -u
mkdir -p gpurun_out
tag=${TAG:-multi}
N=${NGPUS:-4}
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 "${@:3}"; }
: > gpurun_out/${tag}_host_bw.jsonl
for n in 1 2 $N; do
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n tools/host_bw.py 2>> gpurun_out/${tag}.err | grep '^{' >> gpurun_out/${tag}_host_bw.jsonl
done
HOST_BW_BIND=0 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/host_bw.py 2>> gpurun_out/${tag}.err | grep '^{' >> gpurun_out/${tag}_host_bw.jsonl
cat gpurun_out/${tag}_host_bw.jsonl
for n in $N; do
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n --steps 60 --warmup 5 2>> gpurun_out/${tag}.err | grep '^{' > gpurun_out/${tag}_bench_n$n.json
    python -c "import json;d=json.load(open('gpurun_out/${tag}_bench_n$n.json'));print('N',d['n_gpus'],'value',round(d['value']),'e2e',round(d['e2e']['value']),d['e2e']['numa'])"
done

#!/usr/bin/env bash
# gpurun --gpus 8: the kernel-free copy ceiling, the 4K / 8-level line (BASELINE.json configs[4]) and the 1080p line at N = 8.
set -u
mkdir -p gpurun_out
tag=${TAG:-n8}
tr() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 tools/host_bw.py 2>> gpurun_out/${tag}.err | grep '^{' > gpurun_out/${tag}_host_bw.jsonl
cat gpurun_out/${tag}_host_bw.jsonl
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 8 --workload 4k8 --steps 40 --warmup 5 2>> gpurun_out/${tag}.err | grep '^{' > gpurun_out/${tag}_bench_4k8_n8.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 40 --warmup 5 2>> gpurun_out/${tag}.err | grep '^{' > gpurun_out/${tag}_bench_1080p_n8.json
for f in 4k8 1080p; do python -c "import json;d=json.load(open('gpurun_out/${tag}_bench_${f}_n8.json'));print(d['metric'],'N',d['n_gpus'],'lanes',d['config']['lanes_per_gpu'],'value',round(d['value']),'e2e',round(d['e2e']['value']),'level1 frac',round(d['roofline']['fused_level_kernel']['frac'],3))"; done

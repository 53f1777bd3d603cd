#!/usr/bin/env bash
# A/B of whole-step throughput (device-resident and e2e) for library options that the per-kernel profile cannot see
# (lane_groups: concurrent launch chains).  Usage: gpurun -- 'bash tools/gpu_groups.sh "lane_groups=1" "lane_groups=2" ...'
set -u
mkdir -p gpurun_out
tag=${TAG:-groups}
: > gpurun_out/${tag}.jsonl
for o in "$@"; do
    args=""
    for kv in $o; do args="$args --opt $kv"; done
    timeout 300 python bench.py --steps ${STEPS:-100} --warmup 5 --no-cpu-baseline $args > gpurun_out/${tag}_one.json 2>> gpurun_out/${tag}.err
    python - "$o" <<'PY' >> gpurun_out/${tag}.jsonl
import json, sys
try:
    d = json.load(open("gpurun_out/%s_one.json" % __import__("os").environ.get("TAG", "groups")))
    print(json.dumps({"opt": sys.argv[1], "fps": round(d["value"], 1), "ms_per_step": round(d["ms_per_step"], 4), "e2e_fps": round(d["e2e"]["value"], 1)}))
except Exception as e:
    print(json.dumps({"opt": sys.argv[1], "failed": str(e)}))
PY
done
cat gpurun_out/${tag}.jsonl

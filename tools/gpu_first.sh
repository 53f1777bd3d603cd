#!/usr/bin/env bash
# First GPU call of a round: does the whole -m gpu suite pass on hardware (no -x, experimental options included),
# and what do the off-by-default kernel options measure?  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
log() { echo "[gpu_first] $*" | tee -a gpurun_out/gpu_first.log; }
export MC_REQUIRE_REF=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/nvsmi.txt 2>&1
log "1. full GPU suite, no -x, MC_EXPERIMENTAL=1"
MC_EXPERIMENTAL=1 timeout 1200 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1; log "   rc=$?"
tail -5 gpurun_out/pytest_gpu_all.log
log "2. A/B probe of the kernel options (8 and 32 lanes)"
timeout 300 python tests/tools/quick_gpu_probe.py --ab 8,32 > gpurun_out/ab_probe.json 2> gpurun_out/ab_probe.err; log "   rc=$?"
for clip in smooth noise; do
    PROBE_CLIP=$clip PROBE_LANES=8 timeout 120 python tests/tools/quick_gpu_probe.py > gpurun_out/probe_$clip.json 2> gpurun_out/probe_$clip.err; log "   clip $clip rc=$?"
done
log "3. bench (N=1) ours"
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; log "   ours rc=$?"
log "done"

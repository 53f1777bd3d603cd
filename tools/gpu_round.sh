#!/usr/bin/env bash
# One gpurun call that collects everything a round needs from a B200, in order of importance, each step under its own
# timeout so that a surprise in one does not eat the budget of the next.  Everything lands in gpurun_out/.
#
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh'           # full pass (~12-15 min of box time)
#   gpurun --timeout 300  -- 'bash tools/gpu_round.sh quick'     # parity probe + option A/B only (~1 min)
set -u
mkdir -p gpurun_out
mode=${1:-full}
log() { echo "[gpu_round] $*" | tee -a gpurun_out/gpu_round.log; }

log "1. torch-free parity probe + A/B of the off-by-default kernel options (8 and 32 lanes)"
timeout 240 python tests/tools/quick_gpu_probe.py --ab 8,32 > gpurun_out/ab_probe.json 2> gpurun_out/ab_probe.err; log "   rc=$?"
for clip in smooth noise; do   # how content-dependent is the exact-LUT ingest kernel?
    PROBE_CLIP=$clip PROBE_LANES=8 timeout 120 python tests/tools/quick_gpu_probe.py > gpurun_out/probe_$clip.json 2> gpurun_out/probe_$clip.err; log "   clip $clip rc=$?"
done
[ "$mode" = quick ] && exit 0

log "2. GPU parity suite (default paths), then the experimental-option tests"
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; log "   rc=$?"
MC_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_laplace.py -q -m gpu -k "prefetch or egress_tma or fused_tail or ingest_compact or option_combinations" \
    > gpurun_out/pytest_experimental.log 2>&1; log "   experimental rc=$?"

log "3. bench (N=1), both arms"
timeout 400 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; log "   reference rc=$?"
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; log "   ours rc=$?"

log "4. ncu launch list of a short bench run (shares, not absolutes)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; log "   rc=$?"

log "5. ncu --set full of the three largest kernels (3 launches each, after warm-up)"
for k in k_ingest_lab k_egress k_level; do
    timeout 500 ncu --set full --clock-control none --import-source on -k regex:$k -s 12 -c 3 -f -o gpurun_out/full_$k \
        python bench.py --steps 4 --warmup 3 --lanes 16 --no-cpu-baseline > gpurun_out/ncu_$k.log 2>&1; log "   $k rc=$?"
done

log "6. randomised parity on the hardware (5 min)"
timeout 300 python tests/tools/fuzz_parity.py --cases 300 --seed 101 --max-size 400 > gpurun_out/fuzz_gpu.log 2>&1; log "   rc=$?"
log "done"

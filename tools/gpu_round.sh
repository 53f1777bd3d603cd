#!/usr/bin/env bash
# One gpurun call that collects the round's evidence from a B200 on the shipped build, in order of importance, each
# step under its own timeout.  Everything lands in gpurun_out/ (copy what is to be judged into profiles/).
#   gpurun --timeout 1800 -- 'bash tools/gpu_round.sh'
set -u
mkdir -p gpurun_out
tag=${TAG:-round}
log() { echo "[gpu_round] $*" | tee -a gpurun_out/${tag}.log; }
export MC_REQUIRE_REF=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/${tag}_nvsmi.txt 2>&1

log "1. whole GPU suite (no -x)"
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/${tag}_pytest_gpu.log 2>&1; log "   rc=$?"
tail -3 gpurun_out/${tag}_pytest_gpu.log

log "2. bench (N=1), both arms"
timeout 500 python bench.py --impl reference --steps 8 --warmup 2 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; log "   reference rc=$?"
timeout 700 python bench.py > gpurun_out/${tag}_bench_n1.json 2> gpurun_out/${tag}_bench_n1.err; log "   ours rc=$?"

log "3. ncu launch list of a short bench run, one launch chain (shares, not absolutes)"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 50 -c 60 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --opt lane_groups=1 > gpurun_out/${tag}_bench_under_ncu.log 2>&1; log "   rc=$?"

log "4. ncu --set full: ingest, strip egress, fused level 1 and 2 (16 lanes, one chain; k_level launches 5 per step, level 1 first)"
cap() {  # name regex skip count
    timeout 400 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c $4 -f -o gpurun_out/${tag}_full_$1 \
        python bench.py --steps 4 --warmup 3 --lanes 16 --no-cpu-baseline --opt lane_groups=1 > gpurun_out/${tag}_ncu_$1.log 2>&1; log "   $1 rc=$?"
}
cap ingest k_ingest_lab 4 1
cap egress k_egress_strip 4 1
cap level1 k_level 20 1
cap level2 k_level 21 1

log "5. other modes: device-resident fps + kernel tables; ncu --set full of the Phase and Color kernels"
for m in phase color laplace4k laplace_gray; do
    timeout 200 python tools/mode_bench.py $m >> gpurun_out/${tag}_other_modes.jsonl 2>> gpurun_out/${tag}_other_modes.err; log "   $m rc=$?"
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_riesz -s 63 -c 21 -f -o gpurun_out/${tag}_full_riesz \
    python tools/mode_bench.py phase --steps 4 > gpurun_out/${tag}_ncu_riesz.log 2>&1; log "   riesz rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'k_minmax|k_sum_minmax|k_pyrup2x|k_select|k_mask_mul' -s 300 -c 8 -f -o gpurun_out/${tag}_full_color \
    python tools/mode_bench.py color --steps 4 > gpurun_out/${tag}_ncu_color.log 2>&1; log "   color rc=$?"

log "6. compute-sanitizer memcheck over the new kernels (strip egress, 256-bit LUT ingest, lane groups, TMA-staged Riesz tiles)"
timeout 420 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_laplace.py tests/test_gpu_riesz.py -q -m gpu -p no:cacheprovider \
    -k "strip_egress or lane_groups or ingest_warps or tma_staged or flat_regions or small" > gpurun_out/${tag}_compute_sanitizer.txt 2>&1; log "   rc=$?"
tail -4 gpurun_out/${tag}_compute_sanitizer.txt

log "7. single-stream latency; hardware fuzz (2 min)"
timeout 200 python tools/latency.py > gpurun_out/${tag}_latency_lanes1.json 2> gpurun_out/${tag}_latency.err; log "   latency rc=$?"
timeout 150 python tests/tools/fuzz_parity.py --cases 120 --seed 202 --max-size 400 --options > gpurun_out/${tag}_fuzz_gpu.log 2>&1; log "   fuzz rc=$?"
log "done"

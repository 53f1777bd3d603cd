#!/usr/bin/env bash
# One short gpurun call for a kernel change: hardware parity of the Laplace path, the per-kernel A/B table of the
# option sets in $PROBE_VARIANTS at 8 and 32 lanes, the two content variants, and (with "ncu" as $1) one
# `ncu --set full` capture of the kernels named in $NCU_KERNELS.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
tag=${TAG:-ab}
log() { echo "[gpu_ab] $*" | tee -a gpurun_out/${tag}.log; }
export MC_REQUIRE_REF=1
log "1. parity + A/B"
timeout 400 python tests/tools/quick_gpu_probe.py --ab ${AB_LANES:-8,32} > gpurun_out/${tag}_probe.json 2> gpurun_out/${tag}_probe.err; log "   rc=$?"
for clip in smooth noise; do
    PROBE_SKIP_PARITY=1 PROBE_CLIP=$clip PROBE_LANES=8 timeout 120 python tests/tools/quick_gpu_probe.py > gpurun_out/${tag}_$clip.json 2>> gpurun_out/${tag}_probe.err; log "   clip $clip rc=$?"
done
log "2. Laplace GPU tests"
timeout 600 python -m pytest ${PYTEST_ARGS:-tests/test_gpu_laplace.py tests/test_gpu_vs_reference.py tests/test_gpu_golden.py} -q -m gpu -p no:cacheprovider > gpurun_out/${tag}_pytest.log 2>&1; log "   rc=$?"
tail -3 gpurun_out/${tag}_pytest.log
if [ "${1:-}" = ncu ]; then
    log "3. ncu --set full (2 launches each after warm-up, 16 lanes)"
    for k in ${NCU_KERNELS:-k_ingest_lab k_egress k_level}; do
        timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -s 8 -c 2 -f -o gpurun_out/${tag}_full_$k \
            python bench.py --steps 4 --warmup 3 --lanes 16 --no-cpu-baseline > gpurun_out/${tag}_ncu_$k.log 2>&1; log "   $k rc=$?"
    done
fi
log "done"

#!/bin/bash
# One GPU-box session (gpurun): GPU test suite, bit-exactness + timing of the deferred-leaf traversal variants, bench lines,
# shadow pass with and without cut-out grates, ncu launch list + full captures of the traversal kernels. Everything lands in gpurun_out/.
set -u
out=gpurun_out
mkdir -p $out
T0=$(date +%s)
log() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a $out/steps.log; }
: > $out/steps.log
log "start $(nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | head -1)"
timeout 120 python __graft_entry__.py smoke > $out/smoke.log 2>&1; log "smoke rc=$? $(tail -1 $out/smoke.log | cut -c1-200)"
timeout 400 python -m pytest tests/test_texture_parity.py -q -m gpu --maxfail=20 -p no:cacheprovider > $out/pytest_texture.log 2>&1; log "texture tests rc=$? $(tail -1 $out/pytest_texture.log)"
timeout 700 python -m pytest tests -q -m gpu -n 4 --maxfail=20 --durations=12 -p no:cacheprovider --deselect tests/test_texture_parity.py > $out/pytest_all.log 2>&1; log "gpu suite rc=$? $(tail -1 $out/pytest_all.log)"
for d in ${PARITY_DEFER:-1}; do
  VKX_PT_DEFER=$d VKX_PT_DEFER_SHADOW=$d timeout 300 python -m pytest tests/test_ddgi_parity.py tests/test_bvh_parity.py -q -m gpu -n 4 -p no:cacheprovider > $out/pytest_defer_$d.log 2>&1; log "defer $d parity rc=$? $(tail -1 $out/pytest_defer_$d.log)"
done
for d in ${SWEEP_DEFER:-0 1 12}; do
  VKX_PT_DEFER=$d VKX_PT_DEFER_SHADOW=$d timeout 120 python tools/profile_step.py 8 2> $out/sweep_defer_$d.err | tail -1 > $out/sweep_defer_$d.txt; log "sweep defer=$d: $(cat $out/sweep_defer_$d.txt | cut -c1-300)"
done
if [ -f vulkanexp_b200/libvkexp_b200_lazy.so ]; then # A/B of a compile-time variant: triangle words loaded lazily (old) vs all at once (new default)
  VKX_LIB_PATH=$PWD/vulkanexp_b200/libvkexp_b200_lazy.so VKX_PT_DEFER=0 VKX_PT_DEFER_SHADOW=12 timeout 120 python tools/profile_step.py 8 2> $out/sweep_lazy.err | tail -1 > $out/sweep_lazy.txt; log "sweep lazy-load build, defer 0/12: $(cat $out/sweep_lazy.txt | cut -c1-300)"
  VKX_PT_DEFER=0 VKX_PT_DEFER_SHADOW=12 timeout 120 python tools/profile_step.py 8 2> $out/sweep_eager.err | tail -1 > $out/sweep_eager.txt; log "sweep eager-load build, defer 0/12: $(cat $out/sweep_eager.txt | cut -c1-300)"
fi
for b in ${SWEEP_SHADE_BLOCKS:-}; do
  VKX_SHADE_BLOCKS_PER_SM=$b timeout 120 python tools/profile_step.py 8 2> $out/sweep_shade_$b.err | tail -1 > $out/sweep_shade_$b.txt; log "sweep shade blocks/SM=$b: $(cat $out/sweep_shade_$b.txt | cut -c1-300)"
done
best=$(python tools/pick_defer.py $out)
log "best: $best"
timeout 300 python bench.py > $out/bench_default.json 2> $out/bench_default.err; log "bench default rc=$? $(cut -c1-160 $out/bench_default.json)"
env $best timeout 300 python bench.py --no-cpu-baseline > $out/bench_best.json 2> $out/bench_best.err; log "bench best rc=$? $(cut -c1-160 $out/bench_best.json)"
timeout 200 python tools/bench_shadow.py > $out/shadow_plain.json 2> $out/shadow_plain.err; log "shadow plain rc=$? $(cut -c1-400 $out/shadow_plain.json)"
VKX_CFG3_ALPHA=1 timeout 200 python tools/bench_shadow.py > $out/shadow_alpha.json 2> $out/shadow_alpha.err; log "shadow alpha rc=$? $(cut -c1-400 $out/shadow_alpha.json)"
if [ "${SKIP_NCU:-0}" = "1" ]; then log "done (ncu skipped)"; exit 0; fi
# ncu: launch list of the bench command and full captures of the two traversal kernels, with the best variant
env $best timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_${TAG:-r01c}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/bench_under_ncu.log 2>&1; log "ncu launch list rc=$?"
for k in ${PROFILE_DDGI:-k_trace_primary k_trace_shadow}; do
  env $best timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -f -o $out/prof_${TAG:-r01c}_$k python tools/profile_step.py 4 > $out/prof_$k.log 2>&1; log "ncu full $k rc=$?"
done
for k in ${PROFILE_SCREEN:-}; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o $out/prof_${TAG:-r01c}_$k python tools/bench_shadow.py > $out/prof_$k.log 2>&1; log "ncu full $k rc=$?"
done
# the textured kernel variants: cfg3 with cut-out grates (k_direct_light<true>, k_reflect_shade<true>)
for k in ${PROFILE_ALPHA:-k_direct_light}; do
  VKX_CFG3_ALPHA=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o $out/prof_${TAG:-r01c}_${k}_alpha python tools/bench_shadow.py > $out/prof_${k}_alpha.log 2>&1; log "ncu full $k alpha rc=$?"
done
log done

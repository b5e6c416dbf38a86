#!/bin/bash
# Round-end GPU session (gpurun, one B200): smoke, GPU test suite, the bench lines of both arms, cfg3 shadow pass, ncu launch list of the
# bench command and `--set full` captures of the DDGI kernels. Outputs in gpurun_out/; tools/summarise_profiles.py <tag> turns them into profiles/.
set -u
TAG=${TAG:-r02}; out=gpurun_out; mkdir -p $out
T0=$(date +%s); log() { echo "[$(( $(date +%s) - T0 ))s] $*" | tee -a $out/${TAG}_steps.log; }; : > $out/${TAG}_steps.log
log "start $(nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv,noheader | head -1)"
timeout 120 python __graft_entry__.py smoke > $out/${TAG}_smoke.log 2>&1; log "smoke rc=$? $(tail -1 $out/${TAG}_smoke.log | cut -c1-200)"
timeout 900 python -m pytest tests -q -m gpu --durations=8 -p no:cacheprovider > $out/${TAG}_gputest.log 2>&1; log "gpu suite rc=$? $(tail -1 $out/${TAG}_gputest.log)"
timeout 400 python bench.py > $out/${TAG}_bench.json 2> $out/${TAG}_bench.err; log "bench rc=$? $(cut -c1-200 $out/${TAG}_bench.json)"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > $out/${TAG}_bench_reference.json 2> $out/${TAG}_bench_reference.err; log "bench reference rc=$? $(cut -c1-200 $out/${TAG}_bench_reference.json)"
timeout 200 python tools/bench_shadow.py > $out/${TAG}_shadow_plain.json 2> $out/${TAG}_shadow_plain.err; log "shadow plain rc=$? $(cut -c1-300 $out/${TAG}_shadow_plain.json)"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_${TAG}.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary --e2e-steps 2 > $out/${TAG}_bench_under_ncu.log 2>&1; log "ncu launch list rc=$?"
for k in k_trace_primary k_shade_front k_shade_miss k_trace_shadow k_blend_tc k_bin_count k_bin_scatter; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o $out/prof_${TAG}_$k python tools/profile_step.py 4 > $out/prof_${TAG}_$k.log 2>&1; log "ncu full $k rc=$?"
done
log done

#!/bin/bash
# GPU session helper: compute-sanitizer memcheck + racecheck (+ initcheck) over tools/sanitize_small.py for every kernel variant the
# library can run. One summary line per run lands in gpurun_out/<tag>_sanitizer.txt (copied to profiles/); full logs next to it.
tag=${1:-r02}
out=gpurun_out; mkdir -p $out
sum=$out/${tag}_sanitizer.txt; : > $sum
run() { # name, tool, env...
  name=$1; tool=$2; shift 2
  log=$out/${tag}_sanitizer_${name}_${tool}.log
  env "$@" timeout 420 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > $log 2>&1
  rc=$?
  echo "$name [$*] $tool: rc=$rc; $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $log | tail -1); $(grep -c 'sanitize pass ok' $log) pass line(s)" | tee -a $sum
}
for tool in memcheck racecheck; do
  run default $tool VKX_PT_POOL=0 VKX_BLEND=tc
  run pool_simt_defer $tool VKX_PT_POOL=1 VKX_BLEND=simt VKX_PT_DEFER=16 VKX_PT_DEFER_SHADOW=12
done
run default initcheck VKX_PT_POOL=0 VKX_BLEND=tc
run default synccheck VKX_PT_POOL=0 VKX_BLEND=tc
cat $sum

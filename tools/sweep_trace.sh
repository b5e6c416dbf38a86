#!/bin/bash
# GPU session helper: times the DDGI update for build / scheduling variants of the traversal kernels (one bench.py run each).
# usage: tools/sweep_trace.sh out.txt   (variants: default library + libvkexp_b200_<tag>.so builds present next to it)
out=${1:-gpurun_out/sweep_trace.txt}; : > $out
for lib in "" _mb10 _mb12; do
  [ -f vulkanexp_b200/libvkexp_b200$lib.so ] || continue
  for mode in "0 12" "-1 12" "-1 -1" "0 0" "-1 0"; do
    set -- $mode
    r=$(VKX_LIB_PATH=$PWD/vulkanexp_b200/libvkexp_b200$lib.so VKX_PT_DEFER=$1 VKX_PT_DEFER_SHADOW=$2 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline --e2e-steps 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('%.4f' % d['ms_per_step'], {k: round(v,4) for k,v in d['kernel_ms'].items()})")
    echo "lib=${lib:-default} primary=$1 shadow=$2 : $r" | tee -a $out
  done
done

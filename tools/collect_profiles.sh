#!/bin/bash
# Run on the GPU box (gpurun): ncu launch list of the bench command + one `--set full` capture per hot kernel.
# Reports land in gpurun_out/; tools/summarise_profiles.py (run anywhere) turns them into the CSV summaries under profiles/.
set -u
tag=${1:-r01}
out=gpurun_out
mkdir -p $out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $out/bench_under_ncu.log 2>&1
for k in k_trace_primary k_shade_front k_shade_miss k_trace_shadow k_blend; do
  ncu --set full --clock-control none --import-source on -k regex:$k\$ -s 2 -c 1 -f -o $out/prof_${tag}_$k python tools/profile_step.py 4 > $out/prof_$k.log 2>&1
done
for k in k_direct_light k_filter_x k_filter_y k_final_gather k_reflect_shade; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 1 -f -o $out/prof_${tag}_$k python tools/bench_shadow.py > $out/prof_$k.log 2>&1
done
ls -la $out/prof_${tag}_*.ncu-rep

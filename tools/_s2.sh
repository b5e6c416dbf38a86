mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_ddgi_parity.py tests/test_facade.py tests/test_scheduler_parity.py -m gpu -q -x > gpurun_out/r02k_gputest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02k_gputest.log); tail -6 gpurun_out/r02k_gputest.log
for b in tc simt; do VKX_BLEND=$b timeout 300 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline --e2e-steps 20 2>gpurun_out/r02k_bench_$b.err > gpurun_out/r02k_bench_$b.json; python -c "
import json;d=json.load(open('gpurun_out/r02k_bench_$b.json'));print('$b', d['ms_per_step'], d['kernel_ms'])"; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_blend_tc -s 3 -c 1 -f -o gpurun_out/prof_r02k_k_blend_tc python tools/profile_step.py 4 > gpurun_out/prof_r02k_k_blend_tc.log 2>&1; echo "ncu rc=$?"

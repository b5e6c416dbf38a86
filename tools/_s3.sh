mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_ddgi_parity.py tests/test_facade.py tests/test_gather_parity.py tests/test_reflection_parity.py -m gpu -q -x > gpurun_out/r02l_gputest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02l_gputest.log); tail -6 gpurun_out/r02l_gputest.log
for lib in "" _smb5 _smb4; do VKX_LIB_PATH=$PWD/vulkanexp_b200/libvkexp_b200$lib.so timeout 300 python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline --e2e-steps 20 2>gpurun_out/r02l_bench$lib.err > gpurun_out/r02l_bench$lib.json; python -c "
import json;d=json.load(open('gpurun_out/r02l_bench$lib.json'));print('lib$lib', d['ms_per_step'], d['kernel_ms'])"; done
for k in k_blend_tc k_shade_front k_trace_primary; do timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/prof_r02l_$k python tools/profile_step.py 4 > gpurun_out/prof_r02l_$k.log 2>&1; echo "ncu $k rc=$?"; done

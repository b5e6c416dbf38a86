mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02j_gputest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02j_gputest.log); tail -5 gpurun_out/r02j_gputest.log
for pool in 0 1; do VKX_PT_POOL=$pool VKX_BLEND=simt python bench.py --steps 10 --warmup 3 --no-secondary --no-cpu-baseline --e2e-steps 20 2>gpurun_out/r02j_bench_pool$pool.err > gpurun_out/r02j_bench_pool$pool.json; python -c "
import json;d=json.load(open('gpurun_out/r02j_bench_pool$pool.json'));print('pool$pool', d['ms_per_step'], d['kernel_ms'])"; done
(VKX_PT_POOL=1 timeout 600 python -m pytest tests/test_ddgi_parity.py tests/test_bvh_parity.py -m gpu -q > gpurun_out/r02j_gputest_pool.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02j_gputest_pool.log); tail -5 gpurun_out/r02j_gputest_pool.log
python tools/build_time.py > gpurun_out/r02j_build_time.txt 2>&1; tail -5 gpurun_out/r02j_build_time.txt

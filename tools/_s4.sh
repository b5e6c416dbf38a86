mkdir -p gpurun_out
for i in 1 2; do timeout 90 python tools/profile_step.py 8 | tail -1 | cut -c1-200; done
VKX_CFG2_TEXTURED=1 timeout 90 python tools/profile_step.py 6 | tail -1 | cut -c1-200
(timeout 600 python -m pytest tests/test_ddgi_parity.py tests/test_texture_parity.py tests/test_facade.py -m gpu -q -x > gpurun_out/r02aq_gputest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02aq_gputest.log); tail -3 gpurun_out/r02aq_gputest.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 --e2e-steps 100 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(d['ms_per_step'], d['kernel_ms'], 'e2e', d['e2e']['ms_per_step'], 'cfg4 1gpu', d['secondary']['cfg4_single_gpu']['ms_per_step'])"

mkdir -p gpurun_out
for i in 1 2; do timeout 90 python tools/profile_step.py 8 | tail -1 | cut -c1-230; done
(timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r02ai_gputest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02ai_gputest.log); tail -4 gpurun_out/r02ai_gputest.log
timeout 300 python bench.py --no-cpu-baseline --steps 20 --e2e-steps 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print(d['ms_per_step'], d['kernel_ms'], 'cfg4 1gpu', d['secondary']['cfg4_single_gpu']['ms_per_step'], 'shadow4k', d['secondary']['shadow_pass_ms'])"

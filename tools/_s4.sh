mkdir -p gpurun_out
for srt in bins radix bins radix; do VKX_SORT=$srt timeout 90 python tools/profile_step.py 8 | tail -1 | cut -c1-230; done
for srt in bins radix bins radix; do VKX_SORT=$srt timeout 300 python bench.py --no-cpu-baseline --no-secondary --steps 20 --e2e-steps 20 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('$srt', d['ms_per_step'], d['kernel_ms'])"; done

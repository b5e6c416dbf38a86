mkdir -p gpurun_out
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1 | cut -c1-200
(timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02ay_gputest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02ay_gputest.log); tail -3 gpurun_out/r02ay_gputest.log
timeout 400 python bench.py > gpurun_out/r02ay_bench.json 2> gpurun_out/r02ay_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/r02ay_bench.json').read().strip().split('\n')[-1]); print(d['ms_per_step'], d['kernel_ms'], 'e2e', d['e2e']['ms_per_step'], d['cpu_baseline']['value'], d['clocks'])"

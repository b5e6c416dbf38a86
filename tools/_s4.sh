mkdir -p gpurun_out
for i in 1 2; do timeout 90 python tools/profile_step.py 8 | tail -1 | cut -c1-230; done
(timeout 600 python -m pytest tests/test_ddgi_parity.py tests/test_gather_parity.py tests/test_reflection_parity.py -m gpu -q -x > gpurun_out/r02aj_gputest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02aj_gputest.log); tail -3 gpurun_out/r02aj_gputest.log
timeout 120 python tools/diag_parity.py 2>&1 | grep -E "^frame|texels fp32" | head -6

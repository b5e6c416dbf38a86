mkdir -p gpurun_out
timeout 90 python tools/profile_step.py 6 | tail -1
(timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r02ab_gputest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02ab_gputest.log); tail -4 gpurun_out/r02ab_gputest.log
timeout 120 python tools/diag_parity.py > gpurun_out/r02ab_diag.log 2>&1; grep -E "^frame|texels fp32" gpurun_out/r02ab_diag.log | head -8

mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_skinning_parity.py tests/test_bvh_parity.py tests/test_ddgi_parity.py -m gpu -q -x > gpurun_out/r02al_gputest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02al_gputest.log); tail -5 gpurun_out/r02al_gputest.log
for i in 1 2; do timeout 90 python tools/profile_step.py 8 | tail -1 | cut -c1-230; done

mkdir -p gpurun_out
for i in 1 2; do timeout 90 python tools/profile_step.py 8 | tail -1 | cut -c1-230; done
(timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/r02ao_gputest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02ao_gputest.log); tail -4 gpurun_out/r02ao_gputest.log
VKX_CFG2_TEXTURED=1 timeout 90 python tools/profile_step.py 6 | tail -1 | cut -c1-230

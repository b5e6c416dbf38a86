mkdir -p gpurun_out
for d in 0; do echo "== diag $d"; VKX_BLEND_DIAG=$d VKX_BLEND_PROFILE=1 timeout 90 python tools/profile_step.py 3 > gpurun_out/r02q_blend_diag$d.txt 2>&1; echo "rc=$?"; grep "blend_tc profile" gpurun_out/r02q_blend_diag$d.txt | tail -13 | cut -c1-75; done
timeout 90 python tools/profile_step.py 5 | tail -1
(timeout 300 python -m pytest tests/test_ddgi_parity.py tests/test_scheduler_parity.py -m gpu -q -x > gpurun_out/r02q_gputest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02q_gputest.log); tail -4 gpurun_out/r02q_gputest.log

mkdir -p gpurun_out
VKX_BLEND_PROFILE=2 timeout 90 python tools/profile_step.py 3 > gpurun_out/r02ac_blend_diag.txt 2>&1; echo "rc=$?"; grep "blend_tc" gpurun_out/r02ac_blend_diag.txt | tail -29 | cut -c1-160 | grep -v "cta  [1-8]"
for lib in "" _smb5 "" _smb5; do VKX_LIB_PATH=$PWD/vulkanexp_b200/libvkexp_b200$lib.so timeout 90 python tools/profile_step.py 8 | tail -1 | cut -c1-230; done
(timeout 300 python -m pytest tests/test_ddgi_parity.py tests/test_scheduler_parity.py tests/test_facade.py -m gpu -q -x > gpurun_out/r02ac_gputest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02ac_gputest.log); tail -4 gpurun_out/r02ac_gputest.log

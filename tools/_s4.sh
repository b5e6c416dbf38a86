mkdir -p gpurun_out
for i in 1 2; do timeout 90 python tools/profile_step.py 8 | tail -1 | cut -c1-200; done
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02at_launches.csv python tools/profile_step.py 3 > /dev/null 2>&1
grep "k_bin_" gpurun_out/r02at_launches.csv | tail -3 | sed "s/.*k_bin_/k_bin_/" | cut -c1-16,120-400 | rev | cut -c1-12 | rev
(timeout 300 python -m pytest tests/test_ddgi_parity.py tests/test_scheduler_parity.py -m gpu -q -x 2>&1 | tail -2)

mkdir -p gpurun_out
timeout 400 python tools/run_cfg5.py --steps 2 --warmup 1 > gpurun_out/r02_cfg5_n1.json 2> gpurun_out/r02_cfg5_n1.err; echo "cfg5 n1 rc=$?"; tail -c 1800 gpurun_out/r02_cfg5_n1.json; tail -3 gpurun_out/r02_cfg5_n1.err
(timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02x_gputest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02x_gputest.log); tail -5 gpurun_out/r02x_gputest.log

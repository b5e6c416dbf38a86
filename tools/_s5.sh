# 8-GPU session: BASELINE configs[4] (cfg5: ~10 M triangles, 128^3 probes x 256 rays, DDGI + 4K shadows) and configs[3] (bench.py --gpus 8: cfg4 strong scaling)
mkdir -p gpurun_out
N=${1:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/run_cfg5.py --steps 3 --warmup 2 > gpurun_out/r02_cfg5_n$N.json 2> gpurun_out/r02_cfg5_n$N.err; echo "cfg5 rc=$?"; tail -c 1500 gpurun_out/r02_cfg5_n$N.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench rc=$?"; tail -c 1200 gpurun_out/r02_bench_n$N.json

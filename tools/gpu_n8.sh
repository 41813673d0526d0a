#!/bin/bash
# N-GPU timing of the row-sharded path (run on the GPU box from the repo root):
#   gpurun --gpus 8 --timeout 900 -- 'N=8 bash tools/gpu_n8.sh'
set -u
N=${N:-8}
TAG=${TAG:-n$N}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $T --master-port 29523 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/${TAG}_bench.json
timeout 200 $T --master-port 29525 tools/profile_step_sharded.py > gpurun_out/${TAG}_profile.txt 2>&1; echo "profile rc=$?"; grep -v "^CPU" gpurun_out/${TAG}_profile.txt | grep "ms/step\|ms$" | head -40
timeout 300 $T --master-port 29526 bench.py --gpus $N --steps 10 --warmup 3 --workload road3d_3d_g128 > gpurun_out/${TAG}_bench_road3d.json 2> gpurun_out/${TAG}_bench_road3d.err; echo "bench road3d rc=$?"; cut -c1-400 gpurun_out/${TAG}_bench_road3d.json; tail -3 gpurun_out/${TAG}_bench_road3d.err
if [ "${PARITY:-0}" = "1" ]; then
  timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -q > gpurun_out/${TAG}_sharded_pytest.log 2>&1; echo "sharded pytest rc=$?"; tail -5 gpurun_out/${TAG}_sharded_pytest.log
fi

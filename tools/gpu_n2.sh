#!/bin/bash
# 2-GPU validation + timing of the row-sharded path (run on the GPU box from the repo root):
#   gpurun --gpus 2 --timeout 900 -- 'bash tools/gpu_n2.sh'
set -u
N=${N:-2}
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_sharded_gpu.py -m gpu -q > gpurun_out/n${N}_sharded_pytest.log 2>&1; echo "sharded pytest rc=$?"; tail -5 gpurun_out/n${N}_sharded_pytest.log
timeout 200 $T --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/n${N}_bench.json 2> gpurun_out/n${N}_bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/n${N}_bench.json
timeout 200 $T --master-port 29514 bench.py --gpus $N --steps 20 --warmup 3 --single-layout > gpurun_out/n${N}_bench_single.json 2> gpurun_out/n${N}_bench_single.err; echo "bench single-layout rc=$?"; cut -c1-400 gpurun_out/n${N}_bench_single.json
timeout 200 $T --master-port 29515 tools/profile_step_sharded.py > gpurun_out/n${N}_profile.txt 2>&1; echo "profile rc=$?"; grep -v "^CPU" gpurun_out/n${N}_profile.txt | head -45
timeout 200 $T --master-port 29516 bench.py --gpus $N --steps 10 --warmup 3 --workload malaria_2d_g256 > gpurun_out/n${N}_bench_malaria.json 2> gpurun_out/n${N}_bench_malaria.err; echo "bench malaria rc=$?"; cut -c1-300 gpurun_out/n${N}_bench_malaria.json; tail -2 gpurun_out/n${N}_bench_malaria.err
WISKI_SYMM_ALLREDUCE=0 timeout 200 $T --master-port 29517 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/n${N}_bench_nccl_allreduce.json 2> gpurun_out/n${N}_bench_nccl_allreduce.err; echo "bench (NCCL all-reduce) rc=$?"; cut -c1-200 gpurun_out/n${N}_bench_nccl_allreduce.json

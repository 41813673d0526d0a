#!/bin/bash
# 8-GPU pass: parity cases for N = 8, C2 bench at 8 and 4 GPUs, profile, C3 bench:  gpurun --gpus 8 --timeout 1200 -- 'bash tools/gpu_n8_final.sh'
set -u
O=gpurun_out
mkdir -p $O
T8="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
T4="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 400 python -m pytest tests/test_sharded_gpu.py -m gpu -q -k "n8" > $O/n8f_sharded_pytest.log 2>&1; echo "sharded pytest (n8) rc=$?"; tail -3 $O/n8f_sharded_pytest.log
timeout 200 $T8 --master-port 29523 bench.py --gpus 8 --steps 20 --warmup 3 > $O/n8f_bench.json 2> $O/n8f_bench.err; echo "bench n8 rc=$?"; grep "^{" $O/n8f_bench.json | cut -c1-200
timeout 200 $T4 --master-port 29524 bench.py --gpus 4 --steps 20 --warmup 3 > $O/n4f_bench.json 2> $O/n4f_bench.err; echo "bench n4 rc=$?"; grep "^{" $O/n4f_bench.json | cut -c1-200
timeout 200 $T8 --master-port 29525 tools/profile_step_sharded.py > $O/n8f_profile.txt 2>&1; echo "profile rc=$?"; grep "ms/step\|ms$" $O/n8f_profile.txt | head -30
timeout 300 $T8 --master-port 29526 bench.py --gpus 8 --steps 10 --warmup 3 --workload road3d_3d_g128 > $O/n8f_bench_road3d.json 2> $O/n8f_bench_road3d.err; echo "bench road3d n8 rc=$?"; grep "^{" $O/n8f_bench_road3d.json | cut -c1-200

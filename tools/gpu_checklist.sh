#!/bin/bash
# GPU validation still owed by round 1 (everything below was written after the round's GPU budget ran out and has
# only been exercised on CPU with mocked kernels / gloo).  Run from the repo root on a B200 box, e.g.
#   gpurun --timeout 600 -- 'bash tools/gpu_checklist.sh one'          (1 GPU)
#   gpurun --gpus 2 --timeout 400 -- 'bash tools/gpu_checklist.sh two' (2 GPUs; also 4 / 8 with N=4 / N=8)
set -u
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node ${N:-2} --master-addr 127.0.0.1"
case "${1:-one}" in
  one)
    # full parity suite (includes the late host-side rewrites: regression wrapper, MLL, stems, fold-in, state_dict)
    timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
    # (tests/test_next_rows_host_cpu.py — fantasies, BoTorch wrapper, classifier — still runs with mocked kernels only:
    #  give it a DEV switch like tests/model_cases.py to run it on the real kernels)
    # 3droad-shaped config on one GPU (batched fold-in of 19 569 initial points, tensor-core axes, q = 8 under capture)
    timeout 300 python bench.py --workload road3d_3d_g128 --steps 10 --no-cpu-baseline | tee gpurun_out/check_road3d.json | cut -c1-400
    timeout 200 python bench.py --no-cpu-baseline | tee gpurun_out/check_n1.json | cut -c1-300
    ;;
  two)
    # default sharded path, then the dual-layout variant (2 exchanges per step instead of 4)
    timeout 90 $T --master-port 29511 tools/sharded_parity.py --n0 48 --steps 5 --graphs --watchdog 60 | cut -c1-300
    timeout 90 $T --master-port 29512 tools/sharded_parity.py --n0 48 --steps 5 --graphs --dual --watchdog 60 | cut -c1-300
    timeout 120 $T --master-port 29513 bench.py --gpus ${N:-2} --steps 20 --warmup 3 | tee gpurun_out/check_n${N:-2}.json | cut -c1-300
    timeout 120 $T --master-port 29514 bench.py --gpus ${N:-2} --steps 20 --warmup 3 --dual-layout | tee gpurun_out/check_n${N:-2}_dual.json | cut -c1-300
    ;;
esac

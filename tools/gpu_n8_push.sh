# As run for profiles/r02_bench_n8_push_*.json / r02_bench_n4_push_dual_layout.json (the dual layout was still opt-in then:
# --dual-layout / --dual are no-ops today, the single layout is --single-layout / --single).
N=8
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $T --master-port 29523 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/n8p_bench.json 2> gpurun_out/n8p_bench.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/n8p_bench.json
timeout 200 $T --master-port 29524 bench.py --gpus $N --steps 20 --warmup 3 --dual-layout > gpurun_out/n8p_bench_dual.json 2> gpurun_out/n8p_bench_dual.err; echo "bench dual rc=$?"; cut -c1-200 gpurun_out/n8p_bench_dual.json
timeout 200 $T --master-port 29525 tools/profile_step_sharded.py --dual > gpurun_out/n8p_profile_dual.txt 2>&1; echo "profile rc=$?"; grep "ms/step\|ms$" gpurun_out/n8p_profile_dual.txt | head -36
N=4
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 200 $T --master-port 29526 bench.py --gpus $N --steps 20 --warmup 3 --dual-layout > gpurun_out/n4p_bench_dual.json 2> gpurun_out/n4p_bench_dual.err; echo "bench n4 dual rc=$?"; cut -c1-200 gpurun_out/n4p_bench_dual.json

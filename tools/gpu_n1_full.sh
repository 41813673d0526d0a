#!/bin/bash
# 1-GPU validation + measurement pass (run on the GPU box from the repo root):  gpurun --timeout 2400 -- 'bash tools/gpu_n1_full.sh'
set -u
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q > $O/n1_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/n1_pytest.log
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/n1_clocks.csv &
SMI=$!
timeout 500 python bench.py --steps 20 --warmup 5 > $O/n1_bench.json 2> $O/n1_bench.err; echo "bench rc=$?"; cut -c1-300 $O/n1_bench.json
kill $SMI
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $O/n1_bench_reference.json 2> $O/n1_bench_reference.err; echo "bench reference rc=$?"; cut -c1-400 $O/n1_bench_reference.json
for w in road3d_3d_g128 malaria_2d_g256 synthetic_1d_g128; do
  timeout 400 python bench.py --workload $w --steps 10 --warmup 3 --cpu-budget 20 > $O/n1_bench_$w.json 2> $O/n1_bench_$w.err; echo "bench $w rc=$?"; cut -c1-200 $O/n1_bench_$w.json
done
timeout 300 python tools/timeline_step.py > $O/n1_timeline.txt 2> $O/n1_timeline.err; echo "timeline rc=$?"
timeout 300 python tools/online_trace.py --device cuda --steps 400 --out $O/n1_online_trace_gpu.json > $O/n1_online_trace_gpu.log 2>&1; echo "trace rc=$?"; cut -c1-700 $O/n1_online_trace_gpu.log | tail -2
# launch list of two eager steps (every kernel with its device time; cold-cache, serialised)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 700 --csv --log-file $O/n1_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graphs --no-cpu-baseline --no-secondary > $O/n1_launches_bench.log 2>&1; echo "ncu launches rc=$?"
# one full capture of the seven big kernels of one step
#timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tc_gemm3x|lowrank_update2|pair_apply_tc|pair_grad_dir_tc" -s 21 -c 7 \
#    -o $O/n1_ncu_step python bench.py --steps 2 --warmup 3 --no-graphs --no-cpu-baseline --no-secondary > $O/n1_ncu_step.log 2>&1; echo "ncu full rc=$?"; tail -2 $O/n1_ncu_step.log

#!/bin/bash
# streaming second pass of the lane-group kernels: tests, then the condensed-MPC set, random
# problems and the reference distribution with single launch / first window + streaming pass
T=${1:-r2j}
timeout 300 python -m pytest tests/test_passes_gpu.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/${T}_tests.log
timeout 120 python tools/lone_warp_latency.py > gpurun_out/${T}_lone.jsonl 2>gpurun_out/${T}.err
for w in 4 6 8 12; do
  QPB200_PASSES=$w,1 timeout 120 python tools/bench_shapes.py --mpc --shapes 40 --schedule passes --steps 10 2>>gpurun_out/${T}.err | sed "s/^{/{\"w0\": $w, /" >> gpurun_out/${T}_mpc.jsonl
done
timeout 120 python tools/bench_shapes.py --mpc --shapes 40 --schedule single --steps 10 >> gpurun_out/${T}_mpc.jsonl
timeout 120 python tools/bench_shapes.py --mpc --shapes 40 --schedule auto --steps 10 >> gpurun_out/${T}_mpc.jsonl
timeout 120 python tools/bench_shapes.py --shapes 24,32,40 --schedule passes --steps 10 >> gpurun_out/${T}_random.jsonl
timeout 120 python tools/bench_shapes.py --shapes 24,32,40 --schedule single --steps 10 >> gpurun_out/${T}_random.jsonl
timeout 120 python tools/bench_shapes.py --shapes 24,32,40 --schedule auto --steps 10 >> gpurun_out/${T}_random.jsonl
timeout 300 python tools/bench_shapes.py --refdist --shapes 24,32,40 --steps 3 >> gpurun_out/${T}_refdist.jsonl
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/${T}_launches_mpc.csv python tools/bench_shapes.py --mpc --shapes 40 --schedule passes --steps 4 > gpurun_out/${T}_ncu.log 2>&1
cat gpurun_out/${T}_tests.log

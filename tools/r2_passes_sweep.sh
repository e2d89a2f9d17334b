#!/bin/bash
# round 2: window policy of the pass scheduler (QPB200_PASSES=w0,grow) on the condensed-MPC set
# and the reference distribution; ncu launch list of one passes solve (per-pass times)
mkdir -p gpurun_out; T=${1:-r2sweep}
for pol in 24,1 16,1 32,1 48,1 24,0 32,0 12,1; do
  echo "policy $pol" >> gpurun_out/${T}_mpc.jsonl
  QPB200_PASSES=$pol timeout 300 python tools/bench_shapes.py --steps 10 --mpc --shapes 40 --schedule passes >> gpurun_out/${T}_mpc.jsonl 2>> gpurun_out/${T}.err
done
for pol in 24,1 24,2 48,1 16,1; do
  echo "policy $pol" >> gpurun_out/${T}_refdist.jsonl
  QPB200_PASSES=$pol timeout 300 python tools/bench_shapes.py --refdist --steps 3 --shapes 16,40 >> gpurun_out/${T}_refdist.jsonl 2>> gpurun_out/${T}.err
done
echo "count 2^20" >> gpurun_out/${T}_refdist.jsonl
timeout 300 python tools/bench_shapes.py --refdist --steps 3 --shapes 16 --counts 16:1048576 >> gpurun_out/${T}_refdist.jsonl 2>> gpurun_out/${T}.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${T}_mpc_launches.csv python tools/bench_shapes.py --steps 2 --mpc --shapes 40 --schedule passes > gpurun_out/${T}_ncu.log 2>&1
cut -c1-300 gpurun_out/${T}_mpc.jsonl

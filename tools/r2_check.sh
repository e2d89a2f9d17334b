#!/bin/bash
# round 2: GPU tests + device-resident shapes (auto / single schedule) + the reference distribution
mkdir -p gpurun_out; T=${1:-r2a}
nvidia-smi -L > gpurun_out/${T}_gpus.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${T}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tests.log
tail -5 gpurun_out/${T}_tests.log
timeout 300 python tools/bench_shapes.py --steps 10 > gpurun_out/${T}_shapes_auto.jsonl 2> gpurun_out/${T}_shapes_auto.err
timeout 300 python tools/bench_shapes.py --steps 10 --schedule single > gpurun_out/${T}_shapes_single.jsonl 2> gpurun_out/${T}_shapes_single.err
timeout 300 python tools/bench_shapes.py --steps 10 --mpc --shapes 40 > gpurun_out/${T}_mpc_auto.jsonl 2> gpurun_out/${T}_mpc_auto.err
timeout 300 python tools/bench_shapes.py --steps 10 --mpc --shapes 40 --schedule single > gpurun_out/${T}_mpc_single.jsonl 2> gpurun_out/${T}_mpc_single.err
timeout 600 python tools/bench_shapes.py --refdist --steps 3 --shapes 4,8,16,40 > gpurun_out/${T}_refdist.jsonl 2> gpurun_out/${T}_refdist.err
cat gpurun_out/${T}_shapes_auto.jsonl gpurun_out/${T}_shapes_single.jsonl gpurun_out/${T}_mpc_auto.jsonl gpurun_out/${T}_mpc_single.jsonl gpurun_out/${T}_refdist.jsonl | cut -c1-400

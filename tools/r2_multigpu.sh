#!/bin/bash
# round 2: the library's own multi-GPU path (qp_options.device_mask) on N GPUs of one box:
# tests, then bench.py under torchrun (weak scaling, e2e with copy ceiling, e2e_device_mask)
N=${1:-2}; T=${2:-r2_mg${N}}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/${T}_gpus.txt
timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_passes_gpu.py -m gpu -q -rs > gpurun_out/${T}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tests.log
tail -4 gpurun_out/${T}_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
tail -2 gpurun_out/${T}_bench.err
cut -c1-400 gpurun_out/${T}_bench.json

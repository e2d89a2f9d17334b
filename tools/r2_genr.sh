#!/bin/bash
# general rows: the register kernel (qp_genr.cuh) against the shared-memory kernel (qp_gen.cuh):
# tests, then device-resident throughput of both (and of the register kernel with twice the
# lanes per problem, QPB200_GENR_WIDE=1)
mkdir -p gpurun_out; T=${1:-r2g}
timeout 600 python -m pytest tests/test_general_gpu.py -q > gpurun_out/${T}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tests.log
tail -15 gpurun_out/${T}_tests.log
QPB200_GENR_WIDE=1 timeout 600 python -m pytest tests/test_general_gpu.py -q -k "register or identity" > gpurun_out/${T}_tests_wide.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tests_wide.log
tail -3 gpurun_out/${T}_tests_wide.log
rm -f gpurun_out/${T}_bench.jsonl
timeout 300 python tools/bench_general.py --shapes 4,8,16,24,32 --kernel reg >> gpurun_out/${T}_bench.jsonl 2>> gpurun_out/${T}_bench.err
echo '"QPB200_GENR_WIDE=1"' >> gpurun_out/${T}_bench.jsonl
QPB200_GENR_WIDE=1 timeout 300 python tools/bench_general.py --shapes 4,8,16 --kernel reg >> gpurun_out/${T}_bench.jsonl 2>> gpurun_out/${T}_bench.err
[ -n "$2" ] && timeout 300 python tools/bench_general.py --shapes 4,8,16,24,32,40 --kernel smem >> gpurun_out/${T}_bench.jsonl 2>> gpurun_out/${T}_bench.err
cut -c1-200 gpurun_out/${T}_bench.jsonl

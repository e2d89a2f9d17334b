#!/bin/bash
# general rows: tests; register kernel, problems by ticket against the static stride; the
# shared-memory kernel the same
mkdir -p gpurun_out; T=${1:-r2gt}
timeout 600 python -m pytest tests/test_general_gpu.py -q > gpurun_out/${T}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tests.log
tail -3 gpurun_out/${T}_tests.log
rm -f gpurun_out/${T}_bench.jsonl
for c in 65536 1000000; do
timeout 300 python tools/bench_general.py --shapes 4,8,16,32 --kernel reg --count $c >> gpurun_out/${T}_bench.jsonl 2>> gpurun_out/${T}_bench.err
echo '"the four lines above: tickets; the next four: QPB200_GENR_TICKET=0"' >> gpurun_out/${T}_bench.jsonl
QPB200_GENR_TICKET=0 timeout 300 python tools/bench_general.py --shapes 4,8,16,32 --kernel reg --count $c >> gpurun_out/${T}_bench.jsonl 2>> gpurun_out/${T}_bench.err
done
timeout 300 python tools/bench_general.py --shapes 16,40,64 --kernel smem >> gpurun_out/${T}_bench.jsonl 2>> gpurun_out/${T}_bench.err
echo '"the three lines above: tickets; the next three: QPB200_GENR_TICKET=0"' >> gpurun_out/${T}_bench.jsonl
QPB200_GENR_TICKET=0 timeout 300 python tools/bench_general.py --shapes 16,40,64 --kernel smem >> gpurun_out/${T}_bench.jsonl 2>> gpurun_out/${T}_bench.err
cut -c1-175 gpurun_out/${T}_bench.jsonl

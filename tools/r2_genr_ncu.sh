#!/bin/bash
# one --set full capture of the register general-rows kernel, raw + source pages; and the
# cost of a solve without its iterations (max_iter = 1)
mkdir -p gpurun_out; T=${1:-r2gn}; N=${2:-16}
C="python tools/bench_general.py --shapes $N --kernel reg --steps 2"
$C > gpurun_out/${T}_plain.jsonl 2> gpurun_out/${T}_plain.err || exit 1
python tools/bench_general.py --shapes 4,8,16,32 --kernel reg --max-iter 1 >> gpurun_out/${T}_plain.jsonl 2>> gpurun_out/${T}_plain.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qp_genr -s 2 -c 1 -f -o gpurun_out/${T} $C > gpurun_out/${T}_ncu.log 2>&1
ncu -i gpurun_out/${T}.ncu-rep --page raw --csv > gpurun_out/${T}_raw.csv
ncu -i gpurun_out/${T}.ncu-rep --page source --csv --print-source sass > gpurun_out/${T}_src.csv
rm -f gpurun_out/${T}.ncu-rep
cut -c1-220 gpurun_out/${T}_plain.jsonl

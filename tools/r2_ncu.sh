#!/bin/bash
# round 2 ncu evidence: launch list of the bench command, one --set full capture of each
# BASELINE shape's kernel (raw page as CSV), launch list of a PASSES solve (per-pass times)
mkdir -p gpurun_out; T=${1:-r2}
B="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --no-shapes --sustained-s 0"
$B > gpurun_out/${T}_ncu_bench_plain.json 2> gpurun_out/${T}_ncu.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${T}_launches_bench_n16.csv $B > gpurun_out/${T}_ncu1.log 2>&1
cap() { # name kernel-regex command...
  local name=$1 re=$2; shift 2
  "$@" > /dev/null 2>> gpurun_out/${T}_ncu.err || return 1
  ncu --set full --clock-control none --import-source on -k regex:$re -s 3 -c 1 -f -o gpurun_out/${T}_$name "$@" > gpurun_out/${T}_ncu_$name.log 2>&1
  ncu -i gpurun_out/${T}_$name.ncu-rep --page raw --csv > gpurun_out/${T}_${name}_ncu_raw.csv
  rm -f gpurun_out/${T}_$name.ncu-rep
}
cap tpq16 qp_tpq $B
cap tpq4  qp_tpq python tools/bench_shapes.py --steps 2 --shapes 4 --schedule single
cap bpq40 qp_bpq python tools/bench_shapes.py --steps 2 --shapes 40 --schedule single
cap gpq64 qp_gpq python tools/bench_shapes.py --steps 2 --shapes 64 --schedule single
cap gpq64cold qp_gpq python tools/bench_shapes.py --steps 2 --shapes 64 --cold --schedule single
python tools/bench_shapes.py --refdist --steps 1 --shapes 16 > /dev/null 2>> gpurun_out/${T}_ncu.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${T}_launches_passes_refdist_n16.csv \
    python tools/bench_shapes.py --refdist --steps 1 --shapes 16 > gpurun_out/${T}_ncu2.log 2>&1
python tools/ncu_summary.py tpq16=gpurun_out/${T}_tpq16_ncu_raw.csv tpq4=gpurun_out/${T}_tpq4_ncu_raw.csv bpq40=gpurun_out/${T}_bpq40_ncu_raw.csv gpq64=gpurun_out/${T}_gpq64_ncu_raw.csv gpq64cold=gpurun_out/${T}_gpq64cold_ncu_raw.csv > gpurun_out/${T}_ncu_summary.json
cat gpurun_out/${T}_ncu_summary.json

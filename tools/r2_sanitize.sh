#!/bin/bash
# round 2: compute-sanitizer over small solves of every kernel family; ncu of the last (latency
# bound) LIST pass of a PASSES solve on the reference distribution
mkdir -p gpurun_out; T=${1:-r2}
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py > gpurun_out/${T}_sanitizer_$tool.log 2>&1
  echo "$tool rc=$?" >> gpurun_out/${T}_sanitizer_$tool.log
  tail -4 gpurun_out/${T}_sanitizer_$tool.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:qp_tpq_kernel -s 10 -c 1 -f -o gpurun_out/${T}_tpq16_list_tail \
    python tools/bench_shapes.py --refdist --steps 1 --shapes 16 > gpurun_out/${T}_ncu_tail.log 2>&1
ncu -i gpurun_out/${T}_tpq16_list_tail.ncu-rep --page raw --csv > gpurun_out/${T}_tpq16_list_tail_ncu_raw.csv
rm -f gpurun_out/${T}_tpq16_list_tail.ncu-rep
python tools/ncu_summary.py tail=gpurun_out/${T}_tpq16_list_tail_ncu_raw.csv

# ncu evidence of the final tpq kernels: launch list, then one --set full capture each of
# tpq<16> (bench.py) and tpq<4> (bench_shapes.py), raw/details pages exported as CSV
set -x
python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/nf_bench.json 2> gpurun_out/nf_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/nf_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/nf_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:qp_tpq -s 3 -c 1 -f -o gpurun_out/nf_tpq16 \
    python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/nf_ncu2.log 2>&1
ncu -i gpurun_out/nf_tpq16.ncu-rep --page raw --csv > gpurun_out/nf_tpq16_raw.csv
ncu -i gpurun_out/nf_tpq16.ncu-rep --page details --csv > gpurun_out/nf_tpq16_details.csv
python tools/bench_shapes.py --steps 2 --shapes 4 > gpurun_out/nf_s4.json 2> gpurun_out/nf_s4.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:qp_tpq -s 3 -c 1 -f -o gpurun_out/nf_tpq4 \
    python tools/bench_shapes.py --steps 2 --shapes 4 > gpurun_out/nf_ncu3.log 2>&1
ncu -i gpurun_out/nf_tpq4.ncu-rep --page raw --csv > gpurun_out/nf_tpq4_raw.csv
# all shapes with the final code
python tools/bench_shapes.py --steps 30 --shapes 4,16,40,64 > gpurun_out/nf_shapes_main.jsonl 2> gpurun_out/nf_shapes_main.err
python tools/bench_shapes.py --steps 30 --shapes 8,12 > gpurun_out/nf_shapes_other.jsonl 2>> gpurun_out/nf_shapes_main.err

# ncu --set full captures of the final bpq<40> and gpq<64> kernels (raw page as CSV)
set -x
python tools/bench_shapes.py --steps 2 --shapes 40,64 > gpurun_out/ng_s.json 2> gpurun_out/ng_s.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:qp_bpq -s 3 -c 1 -f -o gpurun_out/ng_bpq40 \
    python tools/bench_shapes.py --steps 2 --shapes 40 > gpurun_out/ng_ncu1.log 2>&1
ncu -i gpurun_out/ng_bpq40.ncu-rep --page raw --csv > gpurun_out/ng_bpq40_raw.csv
ncu --set full --clock-control none --import-source on -k regex:qp_gpq -s 3 -c 1 -f -o gpurun_out/ng_gpq64 \
    python tools/bench_shapes.py --steps 2 --shapes 64 > gpurun_out/ng_ncu2.log 2>&1
ncu -i gpurun_out/ng_gpq64.ncu-rep --page raw --csv > gpurun_out/ng_gpq64_raw.csv
rm -f gpurun_out/ng_bpq40.ncu-rep gpurun_out/ng_gpq64.ncu-rep

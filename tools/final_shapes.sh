# every shape of profiles/r1_shapes_all.jsonl with the final kernels, then the bench lines
O=gpurun_out/final_shapes.jsonl; : > $O
B="timeout 200 python tools/bench_shapes.py --steps 30"
$B --shapes 4,8,12,16 --counts 8:1048576,12:1048576 >> $O
$B --shapes 24,32,40 --counts 24:262144,32:262144 >> $O
$B --shapes 40 --mpc >> $O
$B --shapes 48,56 --counts 48:131072,56:131072 >> $O
$B --shapes 64 >> $O
$B --shapes 64 --cold >> $O
$B --shapes 96,128 --counts 96:32768,128:16384 >> $O
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
python bench.py --impl reference > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err

# round-end check as the driver runs it: GPU tests, smoke, both bench arms
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
python bench.py --impl reference > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err
QPB200_PACK_NT=1 python bench.py --no-cpu --steps 20 --e2e-steps 20 > gpurun_out/final_bench_nt.json 2> gpurun_out/final_bench_nt.err
python bench.py --no-cpu --steps 20 --e2e-steps 20 > gpurun_out/final_bench_cached.json 2> gpurun_out/final_bench_cached.err

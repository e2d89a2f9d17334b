# e2e of the host path with the AVX-512 packing against the 128-bit one (same box)
grep -m1 -o "avx512f" /proc/cpuinfo > gpurun_out/pa_cpu.txt; nproc >> gpurun_out/pa_cpu.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/pa_cpu.txt
timeout 120 python -m pytest tests/test_abi.py -x -q > gpurun_out/pa_abi.log 2>&1
for i in 1 2; do
QPB200_DEBUG=1 python bench.py --steps 20 --warmup 5 --no-cpu --e2e-steps 20 > gpurun_out/pa_wide_$i.json 2> gpurun_out/pa_wide_$i.err
QPB200_HOST_PACK_SSE=1 python bench.py --steps 20 --warmup 5 --no-cpu --e2e-steps 20 > gpurun_out/pa_sse_$i.json 2> gpurun_out/pa_sse_$i.err
done
timeout 300 python -m pytest tests -m gpu -x -q -k "host_path" > gpurun_out/pa_tests.log 2>&1

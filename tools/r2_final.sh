#!/bin/bash
# round-2 final evidence (one GPU): the round-end check as the driver runs it (GPU tests, smoke,
# both bench arms), the streaming-pass measurements, the launch list of the MPC solve and one
# --set full capture of the first-pass and the streaming kernel
mkdir -p gpurun_out; T=${1:-r2f}
nvidia-smi -L > gpurun_out/${T}_gpus.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${T}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tests.log
tail -3 gpurun_out/${T}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err
timeout 900 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
bash tools/r2_stream.sh ${T}s > /dev/null 2>&1
M="python tools/bench_shapes.py --mpc --shapes 40 --schedule passes --steps 2"
for k in qp_bpq_kernel qp_bpq_stream; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 3 -c 1 -f -o gpurun_out/${T}_$k $M > gpurun_out/${T}_ncu_$k.log 2>&1
  ncu -i gpurun_out/${T}_$k.ncu-rep --page raw --csv > gpurun_out/${T}_${k}_ncu_raw.csv
  rm -f gpurun_out/${T}_$k.ncu-rep
done
cut -c1-300 gpurun_out/${T}_bench.json

# multi-rank bench lines as the driver launches them (final kernels)
nproc > gpurun_out/mg8_nproc.txt
for N in 8 4 2; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"
$TR bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/mg${N}_final.json 2> gpurun_out/mg${N}_final.err
done

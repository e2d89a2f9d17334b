nproc > gpurun_out/mg_nproc.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
for i in 1 2; do
$TR bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu --e2e-steps 20 > gpurun_out/mg2_plain_$i.json 2> gpurun_out/mg2_plain_$i.err
QPB200_HOST_PACK=2 $TR bench.py --gpus 2 --steps 50 --warmup 5 --no-cpu --e2e-steps 20 > gpurun_out/mg2_mix_$i.json 2> gpurun_out/mg2_mix_$i.err
done

#!/bin/bash
# round 2: end-to-end host path on 8 GPUs of one box: copy ceiling (all ranks at once) and chunk size
N=${1:-8}; T=${2:-r2_mg${N}b}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/${T}_topo.txt 2>&1
nproc >> gpurun_out/${T}_topo.txt; lscpu | grep -i "numa\|model name\|socket" >> gpurun_out/${T}_topo.txt
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N "$@"; }
run --steps 100 --warmup 5 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
for c in 8192 65536 262144; do
  run --steps 20 --warmup 3 --no-shapes --sustained-s 0 --no-device-mask --no-cpu --e2e-chunk $c > gpurun_out/${T}_chunk$c.json 2>> gpurun_out/${T}_bench.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${T}_*.json")):
    try:
        l=json.loads(open(f).read().strip().split("\n")[-1]); e=l["e2e"]
        print(f, "e2e %.1f M  ceiling %.1f M  frac %.2f  upper %.1f M" % (e["value"]/1e6, e["copy_ceiling"]/1e6, e["frac_of_ceiling"], (e.get("caller_packed_upper_layout") or {}).get("value",0)/1e6), "mask", (l.get("e2e_device_mask") or {}).get("value"))
    except Exception as ex: print(f, ex)
PY

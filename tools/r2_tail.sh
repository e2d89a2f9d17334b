#!/bin/bash
# does the static stride of the lane-group kernels leave a tail?  QP/s against batch size
mkdir -p gpurun_out; T=${1:-r2tail}; rm -f gpurun_out/${T}.jsonl
for c in 65536 262144 1048576; do
  timeout 300 python tools/bench_shapes.py --shapes 40 --schedule single --counts 40:$c --steps 5 >> gpurun_out/${T}.jsonl 2>> gpurun_out/${T}.err
done
for c in 65536 262144; do
  timeout 300 python tools/bench_shapes.py --shapes 64 --cold --schedule single --counts 64:$c --steps 5 >> gpurun_out/${T}.jsonl 2>> gpurun_out/${T}.err
  timeout 300 python tools/bench_shapes.py --shapes 32 --schedule single --counts 32:$c --steps 5 >> gpurun_out/${T}.jsonl 2>> gpurun_out/${T}.err
done
python - <<PY
import json
for l in open('gpurun_out/${T}.jsonl'):
    try: d=json.loads(l)
    except Exception: continue
    print(d.get('shape'), d.get('count'), d.get('kernel'), round(d.get('qp_per_s',0)/1e6,2), round(d.get('ms_per_step',0),3), d.get('mean_iters'))
PY
tail -2 gpurun_out/${T}.err

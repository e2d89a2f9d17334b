#!/bin/bash
# small host batches: the kernel working on the pinned staging buffers (default) against the
# one-copy path (QPB200_ZERO_COPY_BYTES=0); GPU tests with the default; latency / crossover table
mkdir -p gpurun_out; T=${1:-r2zc}
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tests.log
tail -3 gpurun_out/${T}_tests.log
QPB200_ZERO_COPY_BYTES=0 timeout 300 python tools/zero_copy_latency.py --save gpurun_out/${T}_a.npz > gpurun_out/${T}.jsonl 2> gpurun_out/${T}.err
timeout 300 python tools/zero_copy_latency.py --save gpurun_out/${T}_b.npz >> gpurun_out/${T}.jsonl 2>> gpurun_out/${T}.err
python - <<PY
import numpy as np, json
a=np.load('gpurun_out/${T}_a.npz'); b=np.load('gpurun_out/${T}_b.npz')
print('identical results:', all(np.array_equal(a[k], b[k]) for k in a.files), len(a.files))
for l in open('gpurun_out/${T}.jsonl'):
    d=json.loads(l); print(d['n'], d['QPB200_ZERO_COPY_BYTES'], 'qf', round(d['qp_admm_qf_one_call_us']['best'],1), round(d['qp_admm_qf_one_call_us']['median'],1), [(r['count'], round(r['best_us'],1), round(r['median_us'],1)) for r in d['rows']])
PY
rm -f gpurun_out/${T}_a.npz gpurun_out/${T}_b.npz
timeout 600 python tools/latency_crossover.py --shapes 16,40 --max-log2 12 > gpurun_out/${T}_latency_crossover.jsonl 2>> gpurun_out/${T}.err
cut -c1-330 gpurun_out/${T}_latency_crossover.jsonl

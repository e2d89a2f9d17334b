#!/bin/bash
# the host path with pageable caller memory: the e2e legs with and without the copier thread
# (QPB200_OUT_COPIER=0 = the path as it was), twice each, on one box
mkdir -p gpurun_out; T=${1:-r2p}
for r in 1 2; do for v in 1 0; do
QPB200_OUT_COPIER=$v timeout 600 python bench.py --steps 20 --no-shapes --no-cpu --no-shared --sustained-s 0 > gpurun_out/${T}_bench_c${v}_$r.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python -c "
import json
l=json.loads(open('gpurun_out/${T}_bench_c${v}_$r.json').read().strip().splitlines()[-1])
e=l['e2e']; print('copier=$v', 'e2e', round(e['value']/1e6,2), 'ceiling', round(e.get('copy_ceiling')/1e6,2), 'pageable', round(e['pageable']['value']/1e6,2), 'driver pageable copy', round(e['pageable']['driver_pageable_copy']/1e6,2))
" | tee -a gpurun_out/${T}_summary.txt
done; done

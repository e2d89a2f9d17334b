#!/bin/bash
# the host path with pageable caller memory: GPU tests (numpy arrays are pageable), then the e2e legs
mkdir -p gpurun_out; T=${1:-r2p}
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_tests.log 2>&1; echo "rc=$?" >> gpurun_out/${T}_tests.log
tail -4 gpurun_out/${T}_tests.log
timeout 600 python bench.py --steps 20 --no-shapes --no-cpu --no-shared --sustained-s 0 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?"
python - <<'PY'
import json,sys
l=json.loads(open('gpurun_out/'+sys.argv[1]+'_bench.json').read().strip().splitlines()[-1]) if False else None
PY
python -c "
import json
l=json.loads(open('gpurun_out/${T}_bench.json').read().strip().splitlines()[-1])
e=l['e2e']; print('e2e', e['value'], 'ceiling', e.get('copy_ceiling'), 'pageable', e['pageable'])
"

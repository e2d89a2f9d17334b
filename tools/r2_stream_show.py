import json, sys, os
T = sys.argv[1]
for f in ['mpc', 'random']:
    fn = f'gpurun_out/{T}_{f}.jsonl'
    if not os.path.exists(fn): continue
    for l in open(fn):
        d = json.loads(l); print(f, d.get('w0'), d['shape'], d['schedule'], round(d['qp_per_s']/1e6, 1), round(d['ms_per_step'], 3), round(d['mean_iters'], 2), d['not_converged'], 'fp32', round(d['fp32']['frac'], 3))
fn = f'gpurun_out/{T}_refdist.jsonl'
if os.path.exists(fn):
    for l in open(fn):
        d = json.loads(l); print(d['shape'], d['count'], round(d['mean_iters'], 1), round(d['single']['qp_per_s']/1e6, 2), round(d['passes']['qp_per_s']/1e6, 2), round(d['speedup'], 3), d['bit_identical'])

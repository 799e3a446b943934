#!/bin/bash
# r02 call 18: the default bench line once more (cfg5 sweeps with the extra warm-ups and the best-of-3 figure)
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
python -c "
import json
d=json.loads(open('gpurun_out/r02_bench_n1.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline_whole_step']['frac'], 'e2e', d['e2e']['ms_per_step'], d['clocks'])
for k,v in d['secondary'].items(): print(' ', k, v if isinstance(v,str) else {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','ms_per_solve','ms_per_step','ms_per_sweep','ms_per_sweep_best_of_3','roofline_frac','write_roofline_frac','error')})
"
tail -3 gpurun_out/r02_bench_n1.err

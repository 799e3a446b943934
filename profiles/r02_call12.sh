#!/bin/bash
# r02 call 12: full GPU suite (empty-batch launches, fast query paths), the reference's own test-suite against this
# package, bench (both arms), ncu of bmm / Levy
set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/r02_gputests.log 2>&1
tail -6 gpurun_out/r02_gputests.log
( time REFSUITE_LOG=gpurun_out/r02_reference_suite_full.log timeout 1200 python tests/reference_suite.py ) > gpurun_out/r02_reference_suite.log 2>&1
tail -3 gpurun_out/r02_reference_suite.log | cut -c1-1800
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
show() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d['ms_per_step'], d['roofline_whole_step']['frac'], 'e2e', d['e2e']['ms_per_step'], {k:round(v['avg_launch_us'],2) for k,v in d.get('kernels',{}).items()}, d['clocks'])
for k,v in (d.get('secondary') or {}).items(): print(' ', k, v if isinstance(v,str) else {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','ms_per_solve','ms_per_step','ms_per_sweep','roofline_frac','write_roofline_frac')})
print({k:round(v['us_per_query'],1) for k,v in d['secondary'].get('cfg5_batch_sweep',{}).items()})
" $1; }
show gpurun_out/r02_bench_n1.json
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; cut -c1-1200 gpurun_out/r02_bench_reference.json
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'bmm_ga' -c 3 -o gpurun_out/r02c12_k python profiles/kernels_for_ncu.py > gpurun_out/r02c12_ncu.log 2>&1; tail -2 gpurun_out/r02c12_ncu.log
du -sh gpurun_out

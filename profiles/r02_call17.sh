#!/bin/bash
# r02 call 17 (final verification): build check, smoke, full GPU suite, bench (both arms), ncu of the final Levy kernels
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
( time python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/r02_gputests.log 2>&1
tail -4 gpurun_out/r02_gputests.log
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
show() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d['ms_per_step'], d['roofline_whole_step']['frac'], 'e2e', d['e2e']['ms_per_step'], {k:round(v['avg_launch_us'],2) for k,v in d.get('kernels',{}).items()}, d['clocks'], d['parity_check']['max_rel_err'])
for k,v in (d.get('secondary') or {}).items(): print(' ', k, v if isinstance(v,str) else {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','ms_per_solve','ms_per_step','ms_per_sweep','roofline_frac','write_roofline_frac','error')})
" $1; }
show gpurun_out/r02_bench_n1.json
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err; cut -c1-300 gpurun_out/r02_bench_reference.json
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'levy_tile' -c 6 -o gpurun_out/r02c17_k python profiles/kernels_for_ncu.py > gpurun_out/r02c17_ncu.log 2>&1; tail -2 gpurun_out/r02c17_ncu.log
python profiles/levy_probe.py | grep "^{"
du -sh gpurun_out

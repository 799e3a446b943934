#!/bin/bash
# r02 call 8: GPU suite; the reference's own test-suite against this package; compute-sanitizer; bench; ncu extras
set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q -p no:cacheprovider ) > gpurun_out/r02c8_gputests.log 2>&1
tail -5 gpurun_out/r02c8_gputests.log
( time timeout 1500 python tests/reference_suite.py ) > gpurun_out/r02_reference_suite.log 2>&1
tail -4 gpurun_out/r02_reference_suite.log | cut -c1-1500
for tool in memcheck racecheck synccheck; do ( time timeout 900 compute-sanitizer --tool $tool python profiles/sanitize_kernels.py ) > gpurun_out/r02_sanitizer_$tool.log 2>&1; tail -4 gpurun_out/r02_sanitizer_$tool.log; done
python bench.py > gpurun_out/r02c8_bench.json 2> gpurun_out/r02c8_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r02c8_bench.json').read().strip().splitlines()[-1])
print('full', d['ms_per_step'], d['roofline_whole_step']['frac'], d['e2e']['ms_per_step'], {k:round(v['avg_launch_us'],2) for k,v in d['kernels'].items()})
for k,v in d['secondary'].items(): print(k, v if isinstance(v,str) else {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','ms_per_solve','ms_per_step','ms_per_sweep','roofline_frac','write_roofline_frac')})
"
CFG4_T=6 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_cfg4_eager.csv python profiles/cfg4_eager.py > gpurun_out/r02c8_ncu_cfg4.log 2>&1; python profiles/launch_shares.py gpurun_out/r02_launches_cfg4_eager.csv | head -24
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'levy_tile|bmm_ga' -c 6 -o gpurun_out/r02c8_k python profiles/kernels_for_ncu.py > gpurun_out/r02c8_ncu.log 2>&1; tail -2 gpurun_out/r02c8_ncu.log
du -sh gpurun_out

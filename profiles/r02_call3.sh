#!/bin/bash
# r02 third GPU call: whole GPU suite on the new kernels, probe, bench (+secondary), ncu captures.
set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q -p no:cacheprovider ) > gpurun_out/r02c3_gputests.log 2>&1
tail -6 gpurun_out/r02c3_gputests.log
python profiles/kernel_probe.py > gpurun_out/r02c3_probe.log 2>&1; cat gpurun_out/r02c3_probe.log
( time python bench.py --steps 5 --warmup 3 ) > gpurun_out/r02c3_bench.json 2> gpurun_out/r02c3_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r02c3_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline_whole_step']['frac'], d['parity_check'], {k:(round(v['avg_launch_us'],2), round(v['frac'],3)) for k,v in (d.get('kernels') or {}).items()})
print(d['cpu_baseline'])
for k,v in d['secondary'].items(): print(k, v if isinstance(v,str) else {a:b for a,b in v.items() if a!='config'})
"; tail -3 gpurun_out/r02c3_bench.err
TSDE_BENCH_REF_BUDGET_S=40 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02c3_bench_ref.json 2> gpurun_out/r02c3_bench_ref.err; cat gpurun_out/r02c3_bench_ref.json | cut -c1-1500
ncu --set full --import-source on --clock-control none -c 40 -o gpurun_out/r02c3_full python profiles/kernels_for_ncu.py > gpurun_out/r02c3_ncu_full.log 2>&1; tail -2 gpurun_out/r02c3_ncu_full.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 40 -c 600 --csv --log-file gpurun_out/r02c3_launches_bench.csv python bench.py --steps 1 --warmup 1 --no-graph --no-secondary --no-cpu --workload cfg2_small > gpurun_out/r02c3_ncu_launch.log 2>&1; tail -2 gpurun_out/r02c3_ncu_launch.log

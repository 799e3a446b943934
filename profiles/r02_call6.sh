#!/bin/bash
# r02 call 6: GPU suite, probe, default bench x3 (run-to-run spread), ncu full of the two cfg2 kernels + launch list
set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q -p no:cacheprovider ) > gpurun_out/r02c6_gputests.log 2>&1
tail -5 gpurun_out/r02c6_gputests.log
python profiles/kernel_probe.py > gpurun_out/r02c6_probe.log 2>&1; cat gpurun_out/r02c6_probe.log
for i in 1 2 3; do python bench.py --steps 5 --warmup 3 --no-secondary --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('short', round(d['ms_per_step'],3), round(d['roofline_whole_step']['frac'],4), d['clocks'])"; done
for i in 1 2; do python bench.py --steps 20 --warmup 3 > gpurun_out/r02c6_bench_$i.json 2> gpurun_out/r02c6_bench_$i.err; python -c "
import json
d=json.loads(open('gpurun_out/r02c6_bench_$i.json').read().strip().splitlines()[-1])
print('full', d['ms_per_step'], d['roofline_whole_step']['frac'], d['e2e']['ms_per_step'], {k:round(v['avg_launch_us'],2) for k,v in d['kernels'].items()})
for k,v in d['secondary'].items(): print(k, v if isinstance(v,str) else {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','ms_per_solve','ms_per_step','ms_per_sweep','roofline_frac','write_roofline_frac')})
"; done
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'MilsteinSeedOp|MilsteinOp<float|levy_tile|bmm_ga' -c 8 -o gpurun_out/r02c6_k python profiles/kernels_for_ncu.py > gpurun_out/r02c6_ncu.log 2>&1; tail -2 gpurun_out/r02c6_ncu.log
ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 300 --csv --log-file gpurun_out/r02c6_launches_cfg2_eager.csv python bench.py --steps 1 --warmup 1 --no-graph --no-secondary --no-cpu --workload cfg2_b262144 > gpurun_out/r02c6_ncu_launch.log 2>&1; tail -2 gpurun_out/r02c6_ncu_launch.log
du -sh gpurun_out; ls -la gpurun_out | tail -8

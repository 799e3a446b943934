#!/bin/bash
# r02 call 11: GPU suite on the rewritten Levy / bmm kernels; ncu of those; bench (+ nvidia-smi sampler period A/B);
# cfg3 launch list
set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/r02_gputests.log 2>&1
tail -8 gpurun_out/r02_gputests.log
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'levy_tile|bmm_ga' -c 9 -o gpurun_out/r02c11_k python profiles/kernels_for_ncu.py > gpurun_out/r02c11_ncu.log 2>&1; tail -2 gpurun_out/r02c11_ncu.log
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
show() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], d['ms_per_step'], d['roofline_whole_step']['frac'], 'e2e', d['e2e']['ms_per_step'], {k:round(v['avg_launch_us'],2) for k,v in d.get('kernels',{}).items()}, d['clocks'])
for k,v in (d.get('secondary') or {}).items(): print(' ', k, v if isinstance(v,str) else {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','ms_per_solve','ms_per_step','ms_per_sweep','roofline_frac','write_roofline_frac')})
" $1; }
show gpurun_out/r02_bench_n1.json
TSDE_BENCH_SAMPLER_MS=1000 python bench.py --no-secondary --no-cpu > gpurun_out/r02c11_bench_sampler1000.json 2>/dev/null; show gpurun_out/r02c11_bench_sampler1000.json
TSDE_BENCH_SAMPLER_MS=50 python bench.py --no-secondary --no-cpu > gpurun_out/r02c11_bench_sampler50.json 2>/dev/null; show gpurun_out/r02c11_bench_sampler50.json
for c in srk_additive_expand srk_additive; do
  CFG3=$c ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_cfg3_$c.csv python profiles/cfg3_eager.py > gpurun_out/r02c11_cfg3_$c.log 2>&1
  python profiles/launch_shares.py gpurun_out/r02_launches_cfg3_$c.csv | head -16
done
du -sh gpurun_out

#!/bin/bash
# r02 call 5: all GPU tests; pipelined RNG A/B (call-3 build vs current); bench (default) ; ncu full of the two cfg2 kernels
set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q -p no:cacheprovider ) > gpurun_out/r02c5_gputests.log 2>&1
tail -6 gpurun_out/r02c5_gputests.log
TORCHSDE_B200_LIB=$PWD/profiles/_ab/libtorchsde_b200_c3.so python profiles/kernel_probe.py > gpurun_out/r02c5_probe_c3.log 2>&1
python profiles/kernel_probe.py > gpurun_out/r02c5_probe_new.log 2>&1
paste -d'\n' gpurun_out/r02c5_probe_c3.log gpurun_out/r02c5_probe_new.log
run() { echo "== $*"; env "$@" python bench.py --steps 5 --warmup 3 --no-secondary --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), round(d['roofline_whole_step']['frac'],4), d['parity_check']['max_rel_err'], {k:round(v['avg_launch_us'],2) for k,v in d['kernels'].items()})"; }
run TORCHSDE_B200_LIB=$PWD/profiles/_ab/libtorchsde_b200_c3.so
run A=1
run TORCHSDE_B200_LIB=$PWD/profiles/_ab/libtorchsde_b200_c3.so
run A=1
( time python bench.py --steps 5 --warmup 3 ) > gpurun_out/r02c5_bench.json 2> gpurun_out/r02c5_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/r02c5_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['roofline_whole_step']['frac'])
for k,v in d['secondary'].items(): print(k, v if isinstance(v,str) else {a:b for a,b in v.items() if a!='config'})
"
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:'MilsteinSeedOp|MilsteinOp<float' -c 4 -o gpurun_out/r02c5_cfg2k python profiles/kernels_for_ncu.py > gpurun_out/r02c5_ncu.log 2>&1; tail -2 gpurun_out/r02c5_ncu.log
ls -la gpurun_out | tail -5

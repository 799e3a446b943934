#!/bin/bash
# r02 second GPU call: full-size parity tests, kernel A/B (r01 build vs current), new bench.py with secondary block.
set -x
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_fullsize.py -x -q -s -p no:cacheprovider ) > gpurun_out/r02c2_fullsize.log 2>&1
tail -25 gpurun_out/r02c2_fullsize.log
TORCHSDE_B200_LIB=$PWD/profiles/_ab/libtorchsde_b200_r01.so python profiles/kernel_probe.py > gpurun_out/r02c2_probe_r01.log 2>&1
python profiles/kernel_probe.py > gpurun_out/r02c2_probe_new.log 2>&1
paste -d'\n' gpurun_out/r02c2_probe_r01.log gpurun_out/r02c2_probe_new.log
( time python bench.py --steps 5 --warmup 3 ) > gpurun_out/r02c2_bench.json 2> gpurun_out/r02c2_bench.err
cat gpurun_out/r02c2_bench.json; tail -5 gpurun_out/r02c2_bench.err
TORCHSDE_B200_LIB=$PWD/profiles/_ab/libtorchsde_b200_r01.so python bench.py --steps 5 --warmup 3 --no-secondary --no-cpu > gpurun_out/r02c2_bench_r01lib.json 2>&1
python -c "
import json
for f in ('gpurun_out/r02c2_bench_r01lib.json','gpurun_out/r02c2_bench.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['ms_per_step'], d['roofline_whole_step']['frac'], {k:(round(v['avg_launch_us'],2), round(v['frac'],3)) for k,v in (d.get('kernels') or {}).items()})
"
( time python -m pytest tests -m gpu -x -q -p no:cacheprovider --deselect tests/test_gpu_fullsize.py ) > gpurun_out/r02c2_gputests.log 2>&1
tail -5 gpurun_out/r02c2_gputests.log

#!/bin/bash
# r02 call 7 (2 GPUs): the N>=2 bisect, the NCCL correctness tests, a full N=2 bench line, the new adjoint tests
set -x
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_multi.py tests/test_gpu_adjoint.py tests/test_gpu_general_tma.py tests/test_gpu_brownian.py -x -q -p no:cacheprovider ) > gpurun_out/r02c7_tests.log 2>&1
tail -6 gpurun_out/r02c7_tests.log
bash profiles/r02_multi_gpu_bisect.sh > gpurun_out/r02c7_bisect_stdout.log 2>&1
cat gpurun_out/r02_bisect.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02c7_bench_n2.json 2> gpurun_out/r02c7_bench_n2.err
python -c "
import json
d=json.loads(open('gpurun_out/r02c7_bench_n2.json').read().strip().splitlines()[-1])
print('N=2 full', d['ms_per_step'], d['per_rank_ms_per_step'], d['roofline_whole_step']['frac'], d['e2e']['ms_per_step'])
for k,v in d['secondary'].items(): print(k, v if isinstance(v,str) else {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','ms_per_solve','ms_per_step','ms_per_sweep','roofline_frac','write_roofline_frac','collective')})
"; tail -3 gpurun_out/r02c7_bench_n2.err

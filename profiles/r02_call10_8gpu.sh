#!/bin/bash
# r02 call 9 (8 GPUs): the full bench line at N=8 (weak: cfg2 + cfg3 + cfg4 sharded with the grad all-reduce + cfg5),
# N=4, and strong scaling of cfg2 at N=8;  2-GPU NCCL correctness tests on this box as well
set -x
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
show() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'N', d['n_gpus'], d['scaling'], 'value', round(d['value']/1e9,3), 'G  ms', round(d['ms_per_step'],3), 'per-rank', d['per_rank_ms_per_step'], 'e2e ms', round(d['e2e']['ms_per_step'],3), 'roof', round(d['roofline_whole_step']['frac'],4))
for k,v in (d.get('secondary') or {}).items(): print('  ', k, v if isinstance(v,str) else {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','ms_per_solve','ms_per_step','ms_per_sweep','roofline_frac','write_roofline_frac','collective')})
" $1; }
$TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err; show gpurun_out/r02_bench_n8.json
$TR --nproc-per-node 8 --master-port 29543 bench.py --gpus 8 --steps 20 --warmup 3 --strong --no-secondary --no-cpu > gpurun_out/r02_bench_n8_strong.json 2> gpurun_out/r02_bench_n8_strong.err; show gpurun_out/r02_bench_n8_strong.json
python bench.py --steps 20 --warmup 3 --no-secondary --no-cpu > gpurun_out/r02_bench_n1_on8box.json 2>/dev/null; show gpurun_out/r02_bench_n1_on8box.json
( time python -m pytest tests/test_gpu_multi.py -x -q -p no:cacheprovider ) > gpurun_out/r02c9_multi.log 2>&1; tail -3 gpurun_out/r02c9_multi.log
tail -3 gpurun_out/r02_bench_n8.err

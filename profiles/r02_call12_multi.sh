#!/bin/bash
# r02 call 12 (N GPUs, N = $NG, default 2): the full bench line at N (weak: cfg2 + cfg3 + cfg4 sharded with the
# parameter-gradient all-reduce + cfg5), strong scaling of cfg2 at N, and the NCCL correctness tests
set -x
NG=${NG:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
show() { python -c "
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], 'N', d['n_gpus'], d['scaling'], 'value', round(d['value']/1e9,3), 'G  ms', round(d['ms_per_step'],3), 'per-rank', d['per_rank_ms_per_step'], 'e2e ms', round(d['e2e']['ms_per_step'],3), 'roof', round(d['roofline_whole_step']['frac'],4))
for k,v in (d.get('secondary') or {}).items(): print('  ', k, v if isinstance(v,str) else {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a in ('value','ms_per_solve','ms_per_step','ms_per_sweep','roofline_frac','write_roofline_frac','collective')})
" $1; }
$TR --nproc-per-node $NG --master-port 29541 bench.py --gpus $NG --steps 20 --warmup 3 --no-cpu > gpurun_out/r02_bench_n$NG.json 2> gpurun_out/r02_bench_n$NG.err; show gpurun_out/r02_bench_n$NG.json
$TR --nproc-per-node $NG --master-port 29543 bench.py --gpus $NG --steps 20 --warmup 3 --strong --no-secondary --no-cpu > gpurun_out/r02_bench_n${NG}_strong.json 2> gpurun_out/r02_bench_n${NG}_strong.err; show gpurun_out/r02_bench_n${NG}_strong.json
python bench.py --steps 20 --warmup 3 --no-secondary --no-cpu > gpurun_out/r02_bench_n1_onmultibox.json 2>/dev/null; show gpurun_out/r02_bench_n1_onmultibox.json
( time python -m pytest tests/test_gpu_multi.py -x -q -p no:cacheprovider ) > gpurun_out/r02c12_multi.log 2>&1; tail -3 gpurun_out/r02c12_multi.log
tail -3 gpurun_out/r02_bench_n$NG.err
# (rides along) Levy A/B libraries built by profiles/build_ab.sh, if any, and the in-tree build
true
python profiles/levy_probe.py | grep "^{" | tee -a gpurun_out/r02_levy_ab2.log

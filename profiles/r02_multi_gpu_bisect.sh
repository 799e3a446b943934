#!/bin/bash
# every rank of an N>=2 run took 51.1-51.3 ms per cfg2 solve against 48.4 ms at N=1 on the same
# node, same clocks, no collective in the timed region.  Bisect what changes between "python bench.py" and
# "torchrun --nproc-per-node 2 bench.py":   gpurun --gpus 2 -- bash profiles/r02_multi_gpu_bisect.sh
set -x
mkdir -p gpurun_out
OUT=gpurun_out/r02_bisect.log
: > $OUT
line() { python -c "
import json,sys
for ln in sys.stdin.read().strip().splitlines():
    try: d=json.loads(ln)
    except Exception: continue
    print('   ms/solve', round(d['ms_per_step'],3), 'per-rank', d.get('per_rank_ms_per_step'), 'tableau-probe-us', (d.get('kernels') or {}).get('step_milstein',{}).get('avg_launch_us'), 'clocks', d.get('clocks'))
"; }
B="bench.py --steps 10 --warmup 3 --no-secondary --no-cpu"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port 29511"
echo "A  N=1 plain python, GPU0" | tee -a $OUT;                 python $B | line | tee -a $OUT
echo "B  N=1 plain python, OMP_NUM_THREADS=1" | tee -a $OUT;    OMP_NUM_THREADS=1 python $B | line | tee -a $OUT
echo "C  N=1 under torchrun (1 rank, no process group)" | tee -a $OUT;  $TR --nproc-per-node 1 $B --gpus 1 2>/dev/null | line | tee -a $OUT
echo "D  N=2 under torchrun, NCCL group (the driver's launch)" | tee -a $OUT;  $TR --nproc-per-node 2 $B --gpus 2 2>/dev/null | line | tee -a $OUT
echo "E  2 ranks under torchrun, NO process group (independent solves, concurrently)" | tee -a $OUT;  TSDE_BENCH_NO_PG=1 $TR --nproc-per-node 2 $B --gpus 2 2>/dev/null | line | tee -a $OUT
echo "F  two plain python processes concurrently (GPU0, GPU1), no torchrun" | tee -a $OUT
(CUDA_VISIBLE_DEVICES=0 python $B | line > gpurun_out/_f0.txt) & (CUDA_VISIBLE_DEVICES=1 python $B | line > gpurun_out/_f1.txt) & wait; cat gpurun_out/_f0.txt gpurun_out/_f1.txt | tee -a $OUT
echo "G  N=2 torchrun + NCCL_P2P_DISABLE=1" | tee -a $OUT;      NCCL_P2P_DISABLE=1 $TR --nproc-per-node 2 $B --gpus 2 2>/dev/null | line | tee -a $OUT
echo "H  N=1 plain python on GPU1 alone" | tee -a $OUT;         CUDA_VISIBLE_DEVICES=1 python $B | line | tee -a $OUT
echo "I  N=1 plain, both GPUs visible, NCCL group of size 1 forced" | tee -a $OUT
TSDE_FORCE_PG=1 $TR --nproc-per-node 1 $B --gpus 1 2>/dev/null | line | tee -a $OUT
echo "J  N=2 torchrun again" | tee -a $OUT;  $TR --nproc-per-node 2 $B --gpus 2 2>/dev/null | line | tee -a $OUT
nvidia-smi --query-gpu=index,clocks.sm,power.draw,temperature.gpu --format=csv | tee -a $OUT
cat $OUT

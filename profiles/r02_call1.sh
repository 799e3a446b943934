#!/bin/bash
# r02 first GPU call: baseline state of main at round start + evidence the verdict asked for.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r02c1_smi.txt
( time python -m pytest tests -m gpu -x -q -p no:cacheprovider ) > gpurun_out/r02c1_gputests.log 2>&1
tail -3 gpurun_out/r02c1_gputests.log
python bench.py --steps 5 --warmup 3 > gpurun_out/r02c1_bench.json 2> gpurun_out/r02c1_bench.err
cat gpurun_out/r02c1_bench.json
# pipe utilisation of the Milstein seed kernel (verdict weak #3)
ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:MilsteinSeedOp -c 2 \
    --metrics sm__inst_executed_pipe_xu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_lsu.sum,sm__inst_executed.sum,sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active \
    -o gpurun_out/r02c1_seed python profiles/kernels_for_ncu.py > gpurun_out/r02c1_ncu_seed.log 2>&1
tail -3 gpurun_out/r02c1_ncu_seed.log
# two-phase TMA consumer A/B (carried over from r01)
bash profiles/next_round_two_phase.sh > gpurun_out/r02c1_two_phase.log 2>&1
tail -12 gpurun_out/r02c1_two_phase.log

#!/bin/bash
# First GPU action of the next round: time the two-phase TMA consumer against the shipped one.
#   gpurun --timeout 600 -- 'bash profiles/next_round_two_phase.sh'
# Builds happen on the GPU box (nvcc is in the image); results land in gpurun_out/.
set -x
mkdir -p gpurun_out
python profiles/gen_tma_ab.py > gpurun_out/ab_shuffle_tree.log 2>&1
TSDE_NVCC_EXTRA='-DTSDE_TMA_TWO_PHASE=1' python -c "import __graft_entry__ as g; g.build(force=True)"
python profiles/gen_tma_ab.py > gpurun_out/ab_two_phase.log 2>&1
python -m pytest tests/test_gpu_general_tma.py -x -q -m gpu > gpurun_out/two_phase_tests.log 2>&1
python -c "import __graft_entry__ as g; g.build(force=True)"   # restore the default build
grep -h "stage_kb=16\|stage_kb=None\|bit-identical" gpurun_out/ab_shuffle_tree.log gpurun_out/ab_two_phase.log
tail -3 gpurun_out/two_phase_tests.log

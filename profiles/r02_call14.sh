#!/bin/bash
# r02 call 14: A/B of the Levy kernels' register budget and of the packed pair arithmetic (profiles/build_ab.sh), then
# the Brownian GPU tests on the packed build
set -x
mkdir -p gpurun_out
for l in profiles/_ab/libtsde_levy_*.so; do TORCHSDE_B200_LIB=$l python profiles/levy_probe.py; done 2>&1 | grep "^{" | tee gpurun_out/r02_levy_ab.log
python profiles/levy_probe.py | grep "^{" | tee -a gpurun_out/r02_levy_ab.log
for l in profiles/_ab/libtsde_levy_5_4_1.so profiles/_ab/libtsde_levy_5_4_0.so; do LEVY_M=8 TORCHSDE_B200_LIB=$l python profiles/levy_probe.py; done 2>&1 | grep "^{" | tee -a gpurun_out/r02_levy_ab.log
( time TORCHSDE_B200_LIB=profiles/_ab/libtsde_levy_5_4_1.so python -m pytest tests/test_gpu_brownian.py tests/test_gpu_fullsize.py tests/test_gpu_edge_cases.py -q -p no:cacheprovider ) > gpurun_out/r02c14_tests_packed.log 2>&1
tail -5 gpurun_out/r02c14_tests_packed.log

#!/bin/bash
# r02 call 4: new GPU tests (variants/logqp/adaptive backprop/broadcast), overlap A/B on cfg2 and cfg3, CTA cap A/B
set -x
mkdir -p gpurun_out
( time python -m pytest tests/test_gpu_variants.py tests/test_gpu_broadcast_g.py tests/test_gpu_solver.py tests/test_gpu_adjoint.py -x -q -p no:cacheprovider ) > gpurun_out/r02c4_tests.log 2>&1
tail -8 gpurun_out/r02c4_tests.log
run() { echo "== $*"; env "$@" python bench.py --steps 5 --warmup 3 --no-secondary --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), round(d['roofline_whole_step']['frac'],4), d['parity_check']['max_rel_err'])"; }
run TSDE_OVERLAP=0
run TSDE_OVERLAP=1
run TSDE_OVERLAP=1 TSDE_EW_CTAS=3
run TSDE_OVERLAP=1 TSDE_EW_CTAS=2
run TSDE_OVERLAP=0 TSDE_EW_CTAS=3
run TSDE_OVERLAP=0
for wl in cfg3_srk_additive cfg3_srk_additive_expand cfg3_euler_general cfg3_heun_general cfg2_srk cfg2_euler; do for ov in 0 1; do echo "== $wl overlap=$ov"; TSDE_OVERLAP=$ov python bench.py --steps 5 --warmup 3 --no-secondary --no-cpu --workload $wl 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), round(d['roofline_whole_step']['frac'],4), (d['parity_check'] or {}).get('max_rel_err'))"; done; done

#!/bin/bash
# r02 call 21: full GPU suite incl. the no-host-synchronisation tests
set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/r02_gputests.log 2>&1
tail -4 gpurun_out/r02_gputests.log
grep -n "FAILED\|synchroniz" gpurun_out/r02_gputests.log | head -20

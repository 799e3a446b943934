#!/bin/bash
# r02 call 20: full GPU suite + the default bench line after the asynchronous key tensor
set -x
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -p no:cacheprovider ) > gpurun_out/r02_gputests.log 2>&1
tail -4 gpurun_out/r02_gputests.log
bash profiles/r02_call18.sh

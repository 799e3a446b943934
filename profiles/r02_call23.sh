#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_edge_cases.py -m gpu -q -p no:cacheprovider -k "synchronise" 2>&1 | grep -E "AssertionError|passed|failed|FAILED" | cut -c1-700

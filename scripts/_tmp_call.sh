#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests/test_gpu_fuzz.py tests/test_gpu_multigpu.py -x -q 2>&1 | grep -v WARNING | tail -30

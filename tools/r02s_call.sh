#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 40 python -m pytest tests -q -m gpu -k "twoplane or shims_match" > gpurun_out/r02s_pytest.log 2>&1; echo "pytest rc $?: $(tail -1 gpurun_out/r02s_pytest.log)" | tee gpurun_out/r02s_summary.txt

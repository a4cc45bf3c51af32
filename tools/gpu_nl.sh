#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_cuda_neighbors.py tests/test_cuda_parity.py -q -m gpu --timeout=200 -x -k "neighbor or unsorted" > gpurun_out/nl_test.log 2>&1; echo "nl test rc=$?"; tail -25 gpurun_out/nl_test.log | cut -c1-400

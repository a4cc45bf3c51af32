#!/bin/bash
mkdir -p gpurun_out
timeout 150 python -m pytest tests -q -m gpu --timeout=120 -x > gpurun_out/gpu_tests_final.log 2>&1; echo "gpu tests rc=$?"; tail -2 gpurun_out/gpu_tests_final.log | cut -c1-200
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2

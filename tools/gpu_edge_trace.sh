#!/bin/bash
mkdir -p gpurun_out
SPK_B200_LIB=$PWD/tools/build/libspk_trace.so timeout 300 python tools/edge_trace.py > gpurun_out/edge_trace.log 2>&1; echo rc=$?; cat gpurun_out/edge_trace.log | tail -45

#!/bin/bash
# linear kernels: tests, then timings.  gpurun --timeout 900 -- 'bash tools/gpu_lin.sh tag [time]'
tag=${1:-lin}; out=gpurun_out/$tag; mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_backward.py tests/test_gpu_graph.py -m gpu -q > $out/pytest_new.log 2>&1; echo "pytest new rc=$?"; grep -n "AssertionError: \|passed\|failed" $out/pytest_new.log | head -30
if [ "$2" == "time" ]; then
timeout 300 python tools/linear_check.py 3xtf32 --time --bwd > $out/lin_3x.log 2>&1; grep -v " ok$" $out/lin_3x.log | tail -14
fi

#!/bin/bash
# quick parity + kernel timing: gpurun --timeout 600 -- 'bash tools/gpu_quick.sh tag'
tag=${1:-quick}; out=gpurun_out/$tag; mkdir -p $out
timeout 300 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $out/pytest_gpu.log
CNF_B200_MIXCDF_GENERIC=1 timeout 120 python tools/bench_mixcdf.py | tee $out/mix_generic.json
timeout 120 python tools/bench_mixcdf.py | tee $out/mix_pipe.json
CNF_B200_MIXCDF_GENERIC=1 timeout 120 python tools/bench_mixcdf.py --inv | tee $out/mixinv_generic.json
timeout 120 python tools/bench_mixcdf.py --inv | tee $out/mixinv_pipe.json

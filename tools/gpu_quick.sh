#!/bin/bash
# quick parity + kernel timing (+ optional ncu of the pipelined kernel):
#   gpurun --timeout 600 -- 'bash tools/gpu_quick.sh tag [ncu]'
tag=${1:-quick}; out=gpurun_out/$tag; mkdir -p $out
timeout 300 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 $out/pytest_gpu.log
timeout 120 python tools/bench_mixcdf.py | tee $out/mix_pipe.json
timeout 120 python tools/bench_mixcdf.py --inv | tee $out/mixinv_pipe.json
if [ "$2" == "ncu" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mixcdf_pipe -s 5 -c 1 -f -o $out/pipe_fwd python tools/bench_mixcdf.py --reps 4 > $out/ncu.log 2>&1
fi

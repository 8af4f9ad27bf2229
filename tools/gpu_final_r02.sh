#!/bin/bash
# Final validation of the round:  gpurun --timeout 1500 -- 'bash tools/gpu_final_r02.sh r02_final'
# full GPU test suite, smoke(), the bench line exactly as the driver runs it, the reference arm.
tag=${1:-r02_final}; out=gpurun_out/$tag; mkdir -p $out
timeout 600 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest_gpu.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $out/smoke.log
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 3 > $out/bench_n1.json 2> $out/bench_n1.err ) 2> $out/bench_n1.time; echo "bench rc=$?"; tail -3 $out/bench_n1.time
( time timeout 400 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > $out/bench_reference.json 2> $out/bench_reference.err ) 2> $out/bench_reference.time; echo "reference rc=$?"; tail -3 $out/bench_reference.time

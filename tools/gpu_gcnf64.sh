#!/bin/bash
# GraphCNF forward at the 8-GPU shard size (64 molecules), wide tiles vs the automatic N tile; GEMM tests
tag=${1:-gcnf64}; out=gpurun_out/$tag; mkdir -p $out
timeout 300 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_graph.py -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $out/pytest.log
CNF_B200_LINEAR_WIDE_TILES=1 timeout 200 python tools/bench_graphcnf.py --reps 5 > $out/wide.json 2>$out/wide.err
timeout 200 python tools/bench_graphcnf.py --reps 5 > $out/auto.json 2>$out/auto.err
CNF_B200_LINEAR_WIDE_TILES=1 timeout 200 python tools/bench_graph.py > $out/gc_wide.json 2>$out/gc_wide.err
timeout 200 python tools/bench_graph.py > $out/gc_auto.json 2>$out/gc_auto.err
tail -n1 $out/wide.json $out/auto.json $out/gc_wide.json $out/gc_auto.json

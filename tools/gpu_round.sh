#!/bin/bash
tag=${1:-run}; out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; tail -15 $out/pytest_gpu.log
for K in 64 16 10; do
timeout 120 python tools/bench_mixcdf.py --K $K --B 1024 > $out/mixcdf_k$K.json 2>&1; cat $out/mixcdf_k$K.json
timeout 120 python tools/bench_mixcdf.py --K $K --B 1024 --inv > $out/mixcdf_k${K}_inv.json 2>&1; cat $out/mixcdf_k${K}_inv.json
done
CNF_B200_MIXCDF_GENERIC=1 timeout 120 python tools/bench_mixcdf.py --K 8 > $out/mixcdf_k8_generic.json 2>&1; cat $out/mixcdf_k8_generic.json
CNF_B200_MIXCDF_GENERIC=1 timeout 120 python tools/bench_mixcdf.py --K 8 --inv > $out/mixcdf_k8_generic_inv.json 2>&1; cat $out/mixcdf_k8_generic_inv.json
timeout 400 python tools/bench_graphcnf.py > $out/graphcnf.log 2>&1; grep '^{' $out/graphcnf.log
timeout 300 python tools/bench_graph.py --train > $out/graph_train.log 2>&1; grep '^{' $out/graph_train.log

#!/bin/bash
# Round profile set: launch list of the bench + one `ncu --set full` capture per dominant kernel.
#   gpurun --timeout 1500 -- 'bash tools/gpu_profile.sh r01e'
tag=${1:-prof}; out=gpurun_out/$tag; mkdir -p $out
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 python -m pytest tests/test_gpu_backward.py -m gpu -q > $out/pytest_bwd.log 2>&1; tail -2 $out/pytest_bwd.log
timeout 300 python tools/bench_train.py > $out/train.json 2>&1; tail -1 $out/train.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $out/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 400 $NCU -k regex:mixcdf_pipe -s 34 -c 1 -o $out/mixcdf_pipe_fwd python bench.py --steps 2 --warmup 3 --no-cpu > $out/ncu1.log 2>&1; echo "ncu pipe rc=$?"
timeout 400 $NCU -k regex:linear_mixcdf -s 10 -c 1 -o $out/linear_mixcdf python bench.py --steps 2 --warmup 3 --no-cpu > $out/ncu2.log 2>&1; echo "ncu fused rc=$?"
timeout 400 $NCU -k regex:linear_tc_kernel -s 40 -c 1 -o $out/linear_tc_3x python tools/bench_graph.py --reps 1 > $out/ncu3.log 2>&1; echo "ncu gemm rc=$?"
timeout 400 $NCU -k regex:graph_aggregate -s 8 -c 1 -o $out/graph_aggregate python tools/bench_graph.py --reps 1 > $out/ncu4.log 2>&1; echo "ncu agg rc=$?"
timeout 400 $NCU -k regex:edge_aggregate -s 60 -c 1 -o $out/edge_aggregate python tools/bench_graphcnf.py --reps 1 --inv-batch 256 > $out/ncu5.log 2>&1; echo "ncu edge rc=$?"
timeout 400 $NCU -k regex:mixcdf_bwd -s 4 -c 1 -o $out/mixcdf_bwd python tools/bench_train.py --reps 1 > $out/ncu6.log 2>&1; echo "ncu bwd rc=$?"
timeout 400 $NCU -k regex:categ_encode_bwd -s 1 -c 1 -o $out/categ_encode_bwd python tools/bench_train.py --reps 1 > $out/ncu7.log 2>&1; echo "ncu encbwd rc=$?"
ls -la $out | head -30

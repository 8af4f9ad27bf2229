#!/bin/bash
# Late round-1 profile set (kernels changed after tools/gpu_profile.sh ran):  gpurun --timeout 1200 -- 'bash tools/gpu_profile2.sh r01p2'
tag=${1:-prof2}; out=gpurun_out/$tag; mkdir -p $out
export PATH=$PATH:/usr/local/cuda/bin
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $out/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 200 $NCU -k regex:mixcdf_pipe -s 34 -c 1 -o $out/mixcdf_pipe_fwd python bench.py --steps 2 --warmup 3 --no-cpu > $out/ncu1.log 2>&1; echo "ncu pipe rc=$?"
timeout 200 $NCU -k regex:linear_mixcdf -s 10 -c 1 -o $out/linear_mixcdf python bench.py --steps 2 --warmup 3 --no-cpu > $out/ncu2.log 2>&1; echo "ncu fused rc=$?"
timeout 150 $NCU -k regex:mixcdf_kernel -s 3 -c 1 -o $out/mixcdf_generic_k64 python tools/bench_mixcdf.py --K 64 --B 1024 --reps 3 > $out/ncu3.log 2>&1; echo "ncu generic rc=$?"
timeout 150 $NCU -k regex:mixcdf_kernel -s 3 -c 1 -o $out/mixcdf_generic_k64_inv python tools/bench_mixcdf.py --K 64 --B 1024 --reps 3 --inv > $out/ncu4.log 2>&1; echo "ncu generic inv rc=$?"
timeout 200 $NCU -k regex:graph_aggregate_bwd -s 4 -c 1 -o $out/graph_aggregate_bwd python tools/bench_graph.py --train --reps 2 > $out/ncu5.log 2>&1; echo "ncu aggbwd rc=$?"
timeout 200 $NCU -k regex:logistic_logprob_rows -s 3 -c 1 -o $out/logprob_rows python bench.py --steps 2 --warmup 3 --no-cpu > $out/ncu6.log 2>&1; echo "ncu logprob rc=$?"
ls -la $out | head -20

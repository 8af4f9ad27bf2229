#!/bin/bash
# Round-2 (second session) profile set:  gpurun --timeout 1500 -- 'bash tools/gpu_profile_r02b.sh r02p'
# launch list of the bench (headline legs + LM training record) and ncu --set full captures of the kernels added this session.
tag=${1:-r02p}; out=gpurun_out/$tag; mkdir -p $out
export PATH=$PATH:/usr/local/cuda/bin
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-graphs > $out/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 200 $NCU -k regex:categ_encode_bwd_tpt16 -s 2 -c 1 -o $out/categ_encode_bwd_tpt16 python tools/bench_train.py --reps 1 > $out/ncu1.log 2>&1; echo "ncu encode bwd rc=$?"
timeout 200 $NCU -k regex:invconv_bwd16 -s 10 -c 1 -o $out/invconv_bwd16 python tools/bench_train.py --reps 1 > $out/ncu2.log 2>&1; echo "ncu invconv bwd rc=$?"
timeout 200 $NCU -k regex:actnorm_bwd4 -s 10 -c 1 -o $out/actnorm_bwd4 python tools/bench_train.py --reps 1 > $out/ncu3.log 2>&1; echo "ncu actnorm bwd rc=$?"
timeout 200 $NCU -k regex:linear_tc -s 8 -c 1 -o $out/linear_tc_small python tools/bench_linear_bn.py --reps 2 --shape 2432 768 384 > $out/ncu4.log 2>&1; echo "ncu small gemm rc=$?"
timeout 200 $NCU -k regex:mixcdf_pipe -s 20 -c 1 -o $out/mixcdf_pipe_fwd python bench.py --steps 2 --warmup 3 --no-cpu --no-graphs --no-train > $out/ncu5.log 2>&1; echo "ncu pipe fwd rc=$?"
ls -la $out | head -30

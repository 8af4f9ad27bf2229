#!/bin/bash
# Round-2 profile set:  gpurun --timeout 1500 -- 'bash tools/gpu_profile_r02.sh r02a'
# launch list of the LM bench + ncu --set full captures of the kernels VERDICT r01 asked evidence for.
tag=${1:-r02a}; out=gpurun_out/$tag; mkdir -p $out
export PATH=$PATH:/usr/local/cuda/bin
NCU="ncu --set full --clock-control none --import-source on -f"
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu --no-graphs"
python tools/ab_invconv.py --reps 30 > $out/ab_invconv.json 2> $out/ab_invconv.err; echo "ab_invconv rc=$?"; cat $out/ab_invconv.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $out/launches_bench.csv $BENCH > $out/ncu_launches.log 2>&1; echo "launch list rc=$?"
timeout 200 $NCU -k regex:invconv_rows -s 10 -c 1 -o $out/invconv_rows_e2e $BENCH > $out/ncu1.log 2>&1; echo "ncu invconv e2e rc=$?"
timeout 200 $NCU -k regex:invconv_rows -s 4 -c 1 -o $out/invconv_rows_plain python tools/ab_invconv.py --reps 3 --only A_rows_plain > $out/ncu2.log 2>&1; echo "ncu invconv plain rc=$?"
timeout 200 $NCU -k regex:linear_tc -s 4 -c 1 -o $out/invconv_tcgen05_3xtf32 python tools/ab_invconv.py --reps 3 --only B_tcgen05_3xtf32 > $out/ncu3.log 2>&1; echo "ncu invconv tcgen05 rc=$?"
timeout 200 $NCU -k regex:mixcdf_pipe -s 3 -c 1 -o $out/mixcdf_pipe_inv python tools/bench_mixcdf.py --inv --reps 3 > $out/ncu4.log 2>&1; echo "ncu pipe inv rc=$?"
timeout 200 $NCU -k regex:categ_encode_tpt -s 4 -c 1 -o $out/categ_encode_tpt $BENCH > $out/ncu5.log 2>&1; echo "ncu encode rc=$?"
timeout 200 $NCU -k regex:linear_mixcdf -s 10 -c 1 -o $out/linear_mixcdf $BENCH > $out/ncu6.log 2>&1; echo "ncu fused rc=$?"
timeout 300 $NCU -k regex:edge_aggregate_kernel -s 40 -c 1 -o $out/edge_aggregate_sampling python tools/bench_graphcnf.py --reps 1 --fwd-batch 64 --inv-batch 1024 > $out/ncu7.log 2>&1; echo "ncu edge agg rc=$?"
ls -la $out | head -30

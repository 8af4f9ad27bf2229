#!/bin/bash
# One GPU-box visit: smoke, GPU parity tests, bench, ncu launch list and a full capture of the
# dominant kernel.  Usage (from the repo root): gpurun --timeout 900 -- 'bash tools/gpu_check.sh r01a'
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/gpu.txt 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $out/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $out/smoke.log
timeout 600 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $out/pytest_gpu.log
tail -5 $out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
cat $out/bench.json; tail -5 $out/bench.err
if [ "$2" != "noprof" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu > $out/ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:mixcdf -s 34 -c 2 -f -o $out/mixcdf_fwd \
    python bench.py --steps 2 --warmup 3 --no-cpu > $out/ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
ls -la $out

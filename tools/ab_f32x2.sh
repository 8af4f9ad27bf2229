#!/bin/bash
# A/B of the packed-fp32 (FFMA2) mixture math against a scalar build: first `CNF_B200_NVCC_FLAGS=-DCNF_NO_F32X2 python -m categoricalnf_b200.build --force && cp categoricalnf_b200/libcnf_b200.so categoricalnf_b200/libcnf_b200_scalar.so.ab && python -m categoricalnf_b200.build --force`, then run this on the GPU box.
cd categoricalnf_b200; cp libcnf_b200.so /tmp/packed.so; cd ..
run() { for i in 1 2 3; do timeout 100 python tools/bench_mixcdf.py --reps 60 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  fwd  %.4f ms (min %.4f)' % (d['ms_median'], d['ms_min']))"; done
        timeout 100 python tools/bench_mixcdf.py --inv --reps 30 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('  inv  %.4f ms' % d['ms_median'])"
        timeout 100 python tools/bench_fused.py 2>&1 | tail -1
        timeout 200 python tools/bench_bwd.py 2>&1 | tail -2; }
echo "== packed"; run
cp categoricalnf_b200/libcnf_b200_scalar.so.ab categoricalnf_b200/libcnf_b200.so
echo "== scalar"; run
cp /tmp/packed.so categoricalnf_b200/libcnf_b200.so
echo "== packed again"; run

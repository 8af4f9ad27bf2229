#!/bin/bash
# timing-only experiments on linear_mixcdf_kernel (results are WRONG under these flags): rebuild with one flag at a time
for flag in "" "-DCNF_FUSED_UNROLL=2" "-DCNF_EXP_NOBIAS" "-DCNF_EXP_NOBARRIER" "-DCNF_EXP_NOTMEMLD" "-DCNF_EXP_NOBIAS -DCNF_EXP_NOBARRIER -DCNF_EXP_NOTMEMLD"; do
    touch categoricalnf_b200/csrc/linear_mixcdf.cu
    CNF_B200_NVCC_FLAGS="$flag" python -m categoricalnf_b200.build > /tmp/build.log 2>&1 || { echo "build failed for $flag"; tail -5 /tmp/build.log; continue; }
    echo "[$flag] $(timeout 120 python tools/bench_fused.py --H 16 --reps 20)"
done
touch categoricalnf_b200/csrc/linear_mixcdf.cu
python -m categoricalnf_b200.build > /dev/null 2>&1

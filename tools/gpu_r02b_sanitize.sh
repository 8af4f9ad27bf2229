#!/bin/bash
# compute-sanitizer over the tests of the kernels added in the second session of round 2 (categ_encode_bwd_tpt16,
# invconv_bwd16, actnorm_bwd4: csrc/categ_bwd.cu, csrc/elementwise_bwd.cu) and the projection kernel with narrow N tiles:
#   gpurun --timeout 1500 -- 'bash tools/gpu_r02b_sanitize.sh'
export PATH=$PATH:/usr/local/cuda/bin
out=gpurun_out; mkdir -p $out
for tool in memcheck racecheck; do
    start=$(date +%s)
    timeout ${SAN_TIMEOUT:-500} compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 99 \
        python -m pytest tests/test_gpu_backward.py -m gpu -q -x -k "categ_encode_backward or invconv_backward or actnorm_backward or training_step or compact or mixcdf_backward" -p no:cacheprovider \
        > $out/r02b_sanitizer_$tool.log 2>&1
    rc=$?
    echo "== $tool rc=$rc $(( $(date +%s) - start )) s: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/r02b_sanitizer_$tool.log | tail -1) | $(grep -E 'passed|failed' $out/r02b_sanitizer_$tool.log | tail -1)"
done

#!/bin/bash
# compute-sanitizer over the tests of the mbarrier / TMA / tensor-memory kernels (mixcdf_pipe.cu, linear_tc.cu, linear_mixcdf.cu):
#   gpurun --timeout 2400 -- 'bash tools/gpu_r02_sanitize.sh'
# Full-BASELINE-size tests are deselected (the tools slow kernels down 10-100x); every fast path is still selected by the
# smaller cases.  Logs -> gpurun_out/r02_sanitizer_{memcheck,racecheck,synccheck}.log, one summary line each.
export PATH=$PATH:/usr/local/cuda/bin
out=gpurun_out; mkdir -p $out
SEL='not full_size and not matches_generic'
for tool in memcheck racecheck synccheck; do
    start=$(date +%s)
    timeout ${SAN_TIMEOUT:-700} compute-sanitizer --tool $tool --print-limit 20 --error-exitcode 99 \
        python -m pytest tests/test_gpu_fastpaths.py tests/test_gpu_tensorcore.py -m gpu -q -x -k "$SEL" -p no:cacheprovider \
        > $out/r02_sanitizer_$tool.log 2>&1
    rc=$?
    echo "== $tool rc=$rc $(( $(date +%s) - start )) s: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out/r02_sanitizer_$tool.log | tail -1) | $(grep -E 'passed|failed' $out/r02_sanitizer_$tool.log | tail -1)"
done

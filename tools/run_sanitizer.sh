#!/usr/bin/env bash
# compute-sanitizer over the small GPU parity tests (ONE tool per gpurun call, see
# /opt/skills/guides/B200_PROFILING.md: several tools in one call have left the GPU unusable).
#   gpurun --timeout 1500 -- 'bash tools/run_sanitizer.sh memcheck'
#   gpurun --timeout 1500 -- 'bash tools/run_sanitizer.sh racecheck'
# The plain run must pass first (a faulting program under a tool is what wedges the device).
set -u
TOOL="${1:-memcheck}"
SEL="${2:-test_pipeline_small_end_to_end or test_kmeans_tensor_path_ties or test_kmeans_tensor_path_hints or test_gram_tensor_path_matches_fp64 or test_msm_vs_oracle or test_counts_golden or test_eigenvalues_lanczos or test_relabel_compact or test_silhouette or test_vamp or test_tile_product or test_f16_kmajor or test_samples_are_reversible}"
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_gpu_bayes.py -x -q -m gpu -k "$SEL" > gpurun_out/sanitizer_plain.log 2>&1 || { echo "plain run failed; not running $TOOL"; tail -5 gpurun_out/sanitizer_plain.log; exit 1; }
timeout 1200 compute-sanitizer --tool "$TOOL" --target-processes all --error-exitcode 3 --log-file "gpurun_out/sanitizer_${TOOL}.log" \
  python -m pytest tests/test_gpu_parity.py tests/test_gpu_tc.py tests/test_gpu_bayes.py -x -q -m gpu -k "$SEL" > "gpurun_out/sanitizer_${TOOL}_pytest.log" 2>&1
rc=$?
echo "compute-sanitizer $TOOL rc=$rc"
grep -E "ERROR SUMMARY|RACECHECK SUMMARY|error" "gpurun_out/sanitizer_${TOOL}.log" | tail -5
tail -3 "gpurun_out/sanitizer_${TOOL}_pytest.log"
exit $rc

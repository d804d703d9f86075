#!/bin/bash
# compute-sanitizer over the kernels added in round 2 (memcheck on the parity tests, racecheck on the shared-memory heavy ones)
mkdir -p gpurun_out
export CUDA_LAUNCH_BLOCKING=0
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/r2_memcheck.log \
    python -m pytest tests/test_gpu_spatial_mean.py tests/test_gpu_gbox.py tests/test_gpu_distill.py tests/test_gpu_host_feed.py \
    "tests/test_gpu_box_inference.py::test_rows_with_non_finite_values_are_dropped_and_renumbered" \
    "tests/test_gpu_box_inference.py::test_equal_scores_resolve_in_candidate_order" \
    "tests/test_gpu_roi_align.py::test_channels_last_output_edge_cases" "tests/test_gpu_roi_heads.py::test_channels_last_pooling_gives_the_reference_results" \
    -m gpu -q -x > gpurun_out/r2_memcheck_pytest.log 2>&1
echo "memcheck rc=$?"; tail -3 gpurun_out/r2_memcheck_pytest.log; grep -E "ERROR SUMMARY|Invalid|error" gpurun_out/r2_memcheck.log | sort | uniq -c | head
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/r2_racecheck.log \
    python -m pytest "tests/test_gpu_box_inference.py::test_rows_with_non_finite_values_are_dropped_and_renumbered" \
    "tests/test_gpu_spatial_mean.py::test_gradient_layout_and_values" "tests/test_gpu_distill.py" -m gpu -q -x > gpurun_out/r2_racecheck_pytest.log 2>&1
echo "racecheck rc=$?"; tail -3 gpurun_out/r2_racecheck_pytest.log; grep -E "RACECHECK SUMMARY|hazard" gpurun_out/r2_racecheck.log | sort | uniq -c | head

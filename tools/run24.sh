#!/bin/bash
# compute-sanitizer over the ROUND-2 kernels (two-level sort, latency pairing kernel, Poseidon, fused PLONK batch, Pallas build, resident bases,
# chained kernel) through the GPU tests that exercise them at small sizes; the round-1 kernels were covered by tools/sanitizer_target.py
mkdir -p gpurun_out
SEL='tests/test_gpu_sort_path.py::test_sorted_path_two_valued_and_sparse_scalars tests/test_gpu_sort_path.py::test_sorted_path_device_entry_and_host_chunk_pipeline_agree tests/test_gpu_poseidon.py::test_device_permutation_reproduces_the_public_vector tests/test_gpu_poseidon.py::test_native_transcript_order_scalars_and_points tests/test_gpu_poseidon.py::test_compressed_points_from_bytes tests/test_gpu_parity.py::test_latency_kernel_dead_pairs_and_grid_stride tests/test_gpu_multi.py::test_msm_with_resident_bases_matches_plain_msm'
S=$(date +%s)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest $SEL -x -q -m gpu > gpurun_out/sanitizer_memcheck24.log 2>&1; echo "memcheck rc=$? in $(( $(date +%s) - S )) s"
tail -4 gpurun_out/sanitizer_memcheck24.log
S=$(date +%s)
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest $SEL -x -q -m gpu > gpurun_out/sanitizer_racecheck24.log 2>&1; echo "racecheck rc=$? in $(( $(date +%s) - S )) s"
tail -4 gpurun_out/sanitizer_racecheck24.log
S=$(date +%s)
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_plonk_verifier.py tests/test_gpu_pasta.py -x -q -m gpu -k "poseidon or ipa or (pallas and 100)" > gpurun_out/sanitizer_memcheck24b.log 2>&1; echo "memcheck-b rc=$? in $(( $(date +%s) - S )) s"
tail -4 gpurun_out/sanitizer_memcheck24b.log

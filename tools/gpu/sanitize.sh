# compute-sanitizer over the CUDA library on small inputs: memcheck (all kernels), racecheck (the SDF sweeps and the
# capture kernel use shared memory / constant tables), initcheck.   gpurun -- 'bash tools/gpu/sanitize.sh'
set -x
S=/usr/local/cuda/bin/compute-sanitizer
T="tests/test_round2_gpu.py::test_long_tiles_first_changes_no_pixel tests/test_round2_gpu.py::test_render_shard_reassembles_the_frame tests/test_round2_gpu.py::test_option_api tests/test_round2_gpu.py::test_pinned_and_pageable_destinations_agree tests/test_round2_gpu.py::test_options_do_not_change_a_frame tests/test_parity_gpu.py::test_edge_cases tests/test_parity_gpu.py::test_camera_batch_and_shards_equal_single_frames tests/test_parity_gpu.py::test_multi_device_distributed_readback tests/test_sdf_gpu.py::test_gpu_sdf_equals_oracle tests/test_sdf_gpu.py::test_tree_build_replicates_to_every_device tests/test_capture.py tests/test_scenegen.py::test_gpu_generator_equals_host_builder"
for tool in memcheck racecheck initcheck; do
  $S --tool $tool --error-exitcode 99 --log-file gpurun_out/sanitize_$tool.log python -m pytest $T -m gpu -x -q -k "not cube and not icosahedron and not long_slab and not 1280 and not 256-0 or edge_cases or capture or replicates or shape1 or shape3 or option_api or options_do_not" > gpurun_out/sanitize_$tool.out 2>&1
  echo "$tool exit $?"; tail -2 gpurun_out/sanitize_$tool.out; tail -3 gpurun_out/sanitize_$tool.log
done

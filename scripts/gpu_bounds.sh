# the GPU suite on a -DTB2_DEBUG_BOUNDS build: every gather index of the step kernels is range-checked on the device
# (the pool's compute-sanitizer is closed); both the default kernel selection and the forced-pipeline pass
set -x
mkdir -p gpurun_out
export TB2_LIB=build_variants/lib_bounds.so
python -c "from traffic_b200 import native; print('bounds build:', native.lib().tb2_debug_bounds_enabled())"
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_bounds.log; cat gpurun_out/pytest_bounds.log
TB2_FORCE_PIPE=1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py tests/test_gpu_predicates.py tests/test_ensemble.py -m gpu -x -q 2>&1 | tail -3 >> gpurun_out/pytest_bounds.log; tail -3 gpurun_out/pytest_bounds.log

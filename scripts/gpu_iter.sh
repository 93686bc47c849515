# kernel iteration: forced-pipeline parity on the parity / predicate / fuzz tests, then bench each variant library (name=path[:ENV=VAL])
set -x
mkdir -p gpurun_out
TB2_FORCE_PIPE=1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_predicates.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_iter.log; cat gpurun_out/pytest_iter.log
bash scripts/gpu_variants_env.sh "$@"

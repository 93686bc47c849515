# parity of the pipelined kernel on every test world (TB2_FORCE_PIPE=1), then the usual suite, then variant benches
set -x
mkdir -p gpurun_out
TB2_FORCE_PIPE=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu_forcepipe.log; cat gpurun_out/pytest_gpu_forcepipe.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
bash scripts/gpu_variants.sh "$@"

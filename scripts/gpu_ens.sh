# ensemble paths: tests (plain, forced pipeline), then the c5 block of the bench
set -x
mkdir -p gpurun_out
python -m pytest tests/test_ensemble.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
TB2_FORCE_PIPE=1 python -m pytest tests/test_ensemble.py tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-small > gpurun_out/bench_ens.json 2> gpurun_out/bench_ens.err; tail -2 gpurun_out/bench_ens.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_ens.json")); e=d["ensemble"]
print("c4 us/step %.2f"%(d["ms_per_step"]*1e3), "ens value %.4g"%e["value"], "s %.3f"%e["seconds"], e["arrived"], e["mean_steps"], e["reward_sum"])
PY

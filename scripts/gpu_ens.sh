# ensemble paths: tests (plain, forced pipeline, forced pipeline + self-clearing tables), then the c5 block of the bench with
# and without self-clearing tables (the trip statistics must be identical)
set -x
mkdir -p gpurun_out
python -m pytest tests/test_ensemble.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
TB2_FORCE_PIPE=1 python -m pytest tests/test_ensemble.py -m gpu -x -q 2>&1 | tail -3
TB2_FORCE_PIPE=1 TB2_SELF_CLEAR=1 python -m pytest tests/test_ensemble.py tests/test_gpu_parity.py tests/test_gpu_fuzz.py -m gpu -x -q 2>&1 | tail -3
for sc in 1 0; do
TB2_SELF_CLEAR=$sc python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-small > gpurun_out/bench_ens_sc$sc.json 2> gpurun_out/bench_ens_sc$sc.err; tail -2 gpurun_out/bench_ens_sc$sc.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_ens_sc$sc.json")); e=d["ensemble"]
print("self_clear=$sc", "c4 us/step %.2f"%(d["ms_per_step"]*1e3), "ens value %.4g"%e["value"], "s %.3f"%e["seconds"], e["arrived"], e["mean_steps"], e["reward_sum"])
PY
done

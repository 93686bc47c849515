set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python bench.py --steps 300 --warmup 30 > gpurun_out/bench_c4.json 2> gpurun_out/bench_c4.err; tail -3 gpurun_out/bench_c4.err; cut -c1-700 gpurun_out/bench_c4.json
python bench.py --workload c5_ensemble --replicas 1024 --steps 200 --warmup 20 --no-small > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; tail -3 gpurun_out/bench_c5.err; cut -c1-900 gpurun_out/bench_c5.json
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -3 gpurun_out/bench_ref.err; cut -c1-600 gpurun_out/bench_ref.json
nproc; grep -m1 'model name' /proc/cpuinfo

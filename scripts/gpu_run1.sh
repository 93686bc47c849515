set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -5
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/pytest_gpu.log; tail -25 gpurun_out/pytest_gpu.log
python bench.py --steps 300 --warmup 20 > gpurun_out/bench_r1_v3.json 2> gpurun_out/bench_r1_v3.err; cat gpurun_out/bench_r1_v3.json; tail -5 gpurun_out/bench_r1_v3.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-small > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1_v3.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-small > gpurun_out/ncu1.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-small > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_f64 -s 10 -c 3 -o gpurun_out/prof_step_f64_r1_v3 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-small > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
ls -la gpurun_out

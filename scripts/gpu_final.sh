# end-of-round measurement set: tests, default bench (both precisions), launch list + full ncu capture of the step kernel
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err; cut -c1-300 gpurun_out/bench_default.json
python bench.py --precision f32 > gpurun_out/bench_default_f32.json 2> gpurun_out/bench_default_f32.err; cut -c1-300 gpurun_out/bench_default_f32.json
python bench.py --impl reference --steps 50 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cut -c1-300 gpurun_out/bench_reference.json
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-small > gpurun_out/plain_a.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-small > gpurun_out/ncu_a.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-small > gpurun_out/plain_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 10 -c 2 -o gpurun_out/prof_step_final python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-small > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_b.log

# end-of-round evidence set: smoke, GPU suites, default bench (both arms), ncu launch list + full capture of the top kernel
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
TB2_FORCE_PIPE=1 python -m pytest tests/test_gpu_parity.py tests/test_gpu_predicates.py tests/test_gpu_fuzz.py tests/test_ensemble.py -m gpu -x -q 2>&1 | tail -2 > gpurun_out/pytest_gpu_forcepipe.log; cat gpurun_out/pytest_gpu_forcepipe.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_driver_settings.json 2> gpurun_out/bench_driver_settings.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-small --no-ensemble > gpurun_out/plain_a.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-small --no-ensemble > gpurun_out/ncu_a.log 2>&1
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-small --no-ensemble > gpurun_out/plain_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"pipe_kernel" -s 10 -c 2 -o gpurun_out/prof_step_final python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-small --no-ensemble > gpurun_out/ncu_b.log 2>&1
tail -2 gpurun_out/ncu_b.log

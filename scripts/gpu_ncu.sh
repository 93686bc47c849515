# ncu capture of the step kernel on the 1M-car bench (after a plain run of the same command exited 0)
set -x
mkdir -p gpurun_out
TAG=${1:-x}
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-small > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"step_kernel|pipe_kernel" -s 10 -c 2 -o gpurun_out/prof_step_f64_$TAG python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-small > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log

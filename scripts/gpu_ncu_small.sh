# ncu capture of the cluster-resident kernel on the 2k-car world (after a plain run of the same command exited 0)
set -x
mkdir -p gpurun_out
TAG=${1:-x}
python bench.py --workload c2_manhattan_2k --steps 500 --warmup 20 --no-cpu-baseline --no-small > gpurun_out/plain_small_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"resident_kernel" -s 1 -c 1 -o gpurun_out/prof_resident_$TAG python bench.py --workload c2_manhattan_2k --steps 500 --warmup 20 --no-cpu-baseline --no-small > gpurun_out/ncu_small_$TAG.log 2>&1
tail -2 gpurun_out/ncu_small_$TAG.log

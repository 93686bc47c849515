# torchrun bench on N GPUs at the driver's settings: every rank one 1M-car replica (weak), the c5 ensemble block sharded
# over the ranks (strong), then the reference arm (N CPU replicas one after the other)
set -x
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_drv_n$N.json 2> gpurun_out/bench_drv_n$N.err; tail -3 gpurun_out/bench_drv_n$N.err; cut -c1-400 gpurun_out/bench_drv_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; cut -c1-300 gpurun_out/bench_ref_n$N.json

set -x
mkdir -p gpurun_out
N=${1:-2}
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 300 --warmup 30 --no-small > gpurun_out/bench_c4_n$N.json 2> gpurun_out/bench_c4_n$N.err; tail -3 gpurun_out/bench_c4_n$N.err; cut -c1-500 gpurun_out/bench_c4_n$N.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload c5_ensemble --replicas 1024 --steps 200 --warmup 20 --no-small --no-cpu-baseline > gpurun_out/bench_c5_n$N.json 2> gpurun_out/bench_c5_n$N.err; tail -3 gpurun_out/bench_c5_n$N.err; cut -c1-500 gpurun_out/bench_c5_n$N.json; python -c "
import json; d=json.load(open('gpurun_out/bench_c5_n$N.json')); print(d.get('rollout'), d['config'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err; cut -c1-300 gpurun_out/bench_ref_n$N.json

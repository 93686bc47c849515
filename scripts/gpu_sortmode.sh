set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sort_neighbour or f32" 2>&1 | tail -5
for nb in atomic sort; do
python bench.py --neighbor $nb --steps 300 --warmup 30 --no-cpu-baseline --no-small > gpurun_out/bench_nb_$nb.json 2> gpurun_out/bench_nb_$nb.err; tail -2 gpurun_out/bench_nb_$nb.err
python -c "
import json; d=json.load(open('gpurun_out/bench_nb_$nb.json')); print('$nb', 'value %.4g'%d['value'], 'us/step %.2f'%(d['ms_per_step']*1e3), 'launches', d['gpu_launches'])"
done

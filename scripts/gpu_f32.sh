set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python scripts/f32_vs_f64.py c3_sf_50k 1000
python scripts/f32_vs_f64.py c4_city_1m 1000 | tail -2
for p in f64 f32; do
python bench.py --precision $p --steps 300 --warmup 30 --no-cpu-baseline > gpurun_out/bench_$p.json 2> gpurun_out/bench_$p.err; tail -2 gpurun_out/bench_$p.err
python -c "
import json; d=json.load(open('gpurun_out/bench_$p.json')); print('$p', 'value %.4g'%d['value'], 'us/step %.2f'%(d['ms_per_step']*1e3), 'frac %.3f'%d['roofline']['frac'], 'e2e %.4g'%d['e2e']['value'], '2k us/step %.2f'%d['also']['us_per_step'])"
done

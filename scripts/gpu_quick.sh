# quick regression: all GPU tests + default-size bench summary
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 500 --warmup 30 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -2 gpurun_out/bench_quick.err
python -c "
import json; d=json.load(open('gpurun_out/bench_quick.json')); print('value %.4g'%d['value'], 'us/step %.2f'%(d['ms_per_step']*1e3), 'frac %.3f'%d['roofline']['frac'], 'e2e %.4g'%d['e2e']['value'], '2k us/step %.2f'%d['small_world']['us_per_step'], 'facade us/step %.1f'%d['facade']['us_per_step'], 'clocks', d['clocks'])"

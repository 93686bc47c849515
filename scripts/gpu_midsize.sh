set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for w in c3_sf_50k c2_manhattan_2k; do python bench.py --workload $w --steps 2000 --warmup 100 --no-small --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['config']['workload'], 'us/step %.2f'%(d['ms_per_step']*1e3), 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'launches', d['gpu_launches'])"; done
TB2_NO_PERSISTENT=1 python bench.py --workload c3_sf_50k --steps 2000 --warmup 100 --no-small --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('per-step', d['config']['workload'], 'us/step %.2f'%(d['ms_per_step']*1e3), 'value %.4g'%d['value'], 'launches', d['gpu_launches'])"

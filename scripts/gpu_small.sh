# 2k-car world (BASELINE configs[1]): persistent cluster kernel vs one launch per step
set -x
python bench.py --workload c2_manhattan_2k --steps 4000 --warmup 200 --no-cpu-baseline --no-small 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('persistent', 'us/step %.2f'%(d['ms_per_step']*1e3), 'value %.4g'%d['value'], 'launches', d['gpu_launches'])"
TB2_NO_PERSISTENT=1 python bench.py --workload c2_manhattan_2k --steps 4000 --warmup 200 --no-cpu-baseline --no-small 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('per-step launches', 'us/step %.2f'%(d['ms_per_step']*1e3), 'value %.4g'%d['value'], 'launches', d['gpu_launches'])"

set -x
mkdir -p gpurun_out
for so in 0 1; do
  for st in 50 300 1000; do
    TB2_SPATIAL_ORDER=$so python bench.py --steps $st --warmup 3 --no-cpu-baseline --no-small > gpurun_out/sp_${so}_$st.json 2> gpurun_out/sp_${so}_$st.err
    python -c "
import json; d=json.load(open('gpurun_out/sp_${so}_$st.json')); print('spatial=$so steps=$st us/step %.2f value %.4g'%(d['ms_per_step']*1e3, d['value']))"
  done
done

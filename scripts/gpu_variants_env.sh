# like gpu_variants.sh but each variant may carry env settings: name=path[:ENV=VAL[:ENV=VAL]]
set -x
mkdir -p gpurun_out
for kv in "$@"; do
  name="${kv%%=*}"; rest="${kv#*=}"; path="${rest%%:*}"; envs=""
  if [ "$rest" != "$path" ]; then envs="$(echo "${rest#*:}" | tr ':' ' ')"; fi
  env $envs TB2_LIB="$path" python bench.py --steps 300 --warmup 30 --no-cpu-baseline --no-small > gpurun_out/variant_$name.json 2> gpurun_out/variant_$name.err
  python - <<PY
import json
d=json.load(open("gpurun_out/variant_$name.json"))
print("$name", "value %.4g"%d["value"], "us/step %.2f"%(d["ms_per_step"]*1e3), "frac %.3f"%d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"])
PY
done

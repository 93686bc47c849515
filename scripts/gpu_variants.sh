# run bench.py for each kernel variant library given as arguments (name=path), quick settings
set -x
mkdir -p gpurun_out
for kv in "$@"; do
  name="${kv%%=*}"; path="${kv#*=}"
  TB2_LIB="$path" python bench.py --steps 300 --warmup 30 --no-cpu-baseline > gpurun_out/variant_$name.json 2> gpurun_out/variant_$name.err
  python - <<PY
import json
d=json.load(open("gpurun_out/variant_$name.json"))
print("$name", "value %.4g"%d["value"], "us/step %.2f"%(d["ms_per_step"]*1e3), "frac %.3f"%d["roofline"]["frac"], "e2e %.4g"%d["e2e"]["value"], "2k us/step %.2f"%d["small_world"]["us_per_step"])
PY
done

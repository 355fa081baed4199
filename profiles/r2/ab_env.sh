#!/bin/bash
# like ab.sh, for environment settings instead of builds: profiles/r2/ab_env.sh OUT NAME:VAR=VALUE ...
out="$1"; shift
for v in "$@"; do
  name="${v%%:*}"; kv="${v#*:}"
  env $kv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --e2e-steps 5 > "gpurun_out/${out}_${name}.json" 2> "gpurun_out/${out}_${name}.err" || tail -3 "gpurun_out/${out}_${name}.err"
  python -c "
import json,sys; d=json.load(open('gpurun_out/${out}_${name}.json'))
k=d['roofline']['kernel_ms_per_tick']
print('${name}', round(d['value']/1e9,2), round(d['ms_per_step'],4), round(d['roofline']['frac'],3), {a:round(b,4) for a,b in k.items()}, d['state_digest']['sum'])"
done

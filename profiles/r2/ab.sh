#!/bin/bash
# profiles/r2/ab.sh OUT NAME:LIB ... — the driver-protocol bench for several builds of the library on the same box
# (GRIDDY_CUDA_LIB selects the build; "main" = the in-tree one).  One JSON line per build in gpurun_out/OUT_NAME.json.
out="$1"; shift
for v in "$@"; do
  name="${v%%:*}"; lib="${v#*:}"
  if [ "$lib" = "main" ]; then unset GRIDDY_CUDA_LIB; else export GRIDDY_CUDA_LIB="$PWD/$lib"; fi
  python bench.py --steps "${AB_STEPS:-20}" --warmup 5 --no-cpu-baseline --e2e-steps 5 > "gpurun_out/${out}_${name}.json" 2> "gpurun_out/${out}_${name}.err" || tail -3 "gpurun_out/${out}_${name}.err"
  python -c "
import json,sys; d=json.load(open('gpurun_out/${out}_${name}.json'))
k=d['roofline']['kernel_ms_per_tick']
print('${name}', round(d['value']/1e9,2), round(d['ms_per_step'],4), round(d['roofline']['frac'],3), {a:round(b,4) for a,b in k.items()}, d['state_digest']['sum'])"
done

timeout 60 python -m pytest tests -m gpu -x -q -k "golden or positions" 2>&1 | tail -1
timeout 40 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 90 python bench.py > gpurun_out/bench_final3.json 2> gpurun_out/bench_final3.err; grep -i "affinity\|bound to" gpurun_out/bench_final3.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_final3.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],4), round(d["value"]/1e9,2), round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"]/1e9,2), "full", round(d["e2e_full_state"]["value"]/1e9,2), "cpu", round(d["cpu_baseline"]["value"]/1e6,2))
PY

timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 1000 --warmup 20 --no-cpu-baseline --e2e-steps 20 2> gpurun_out/r1s7_mg4.err | grep "^{" > gpurun_out/r1s7_mg4.json; tail -2 gpurun_out/r1s7_mg4.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/r1s7_mg4.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], d["e2e"]["d2h_bytes_per_step"], "full", d["e2e_full_state"]["value"])
print(d["roofline"]["kernel_ms_per_tick"]); print(d["world"]["per_rank"])
PY

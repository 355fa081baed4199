python -X faulthandler -m pytest tests -m gpu -x -q -k "partitioned or c4_one or nccl" > gpurun_out/pytest_r1s10.log 2>&1; tail -3 gpurun_out/pytest_r1s10.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 1000 --warmup 20 --no-cpu-baseline --e2e-steps 30 2> gpurun_out/r1s10_mg2.err | grep "^{" > gpurun_out/r1s10_mg2.json; tail -3 gpurun_out/r1s10_mg2.err | cut -c1-300
python - <<PY
import json
d=json.loads(open("gpurun_out/r1s10_mg2.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], d["e2e"]["d2h_bytes_per_step"], "full", d["e2e_full_state"]["value"])
print(d["roofline"]["kernel_ms_per_tick"])
PY

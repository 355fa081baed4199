export GRIDDY_CUDA_LIB=$PWD/griddy_b200/csrc/variants/libgriddy_cuda_trace.so GRIDDY_DUMP_TRACE=$PWD/gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 4 --steps 1000 --warmup 20 --no-cpu-baseline --e2e-steps 4 --profile-ticks 4 2> gpurun_out/r1s9_mg4.err | grep "^{" > gpurun_out/r1s9_mg4.json; grep trace gpurun_out/r1s9_mg4.err | head -4
python - <<PY
import json
d=json.loads(open("gpurun_out/r1s9_mg4.json").read().strip().splitlines()[-1])
print(d["ms_per_step"], d["value"])
PY

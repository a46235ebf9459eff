#!/bin/bash
# strong-scaling bench line at N GPUs of one box (gpurun --gpus N): python tools/... N
N=${1:-8}
mkdir -p gpurun_out
export PIQMC_WATCHDOG_MS=20000
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "rc=$?"
python - <<PY
import json
lines = [l for l in open("gpurun_out/r2_bench_n$N.json").read().splitlines() if l.startswith("{")]
d = json.loads(lines[-1])
print("N=$N value %.3e e2e %.3e ms/step %.3f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]), {k: round(1e3 * v, 1) for k, v in d["e2e"]["breakdown_s"].items()})
PY
if [ "$2" = weak ]; then
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --steps 20 --warmup 5 --replicas $((4096 * N)) > gpurun_out/r2_bench_n${N}_weak.json 2> gpurun_out/r2_bench_n${N}_weak.err
python - <<PY
import json
lines = [l for l in open("gpurun_out/r2_bench_n${N}_weak.json").read().splitlines() if l.startswith("{")]
d = json.loads(lines[-1])
print("N=$N weak value %.3e e2e %.3e" % (d["value"], d["e2e"]["value"]), d["e2e"].get("overlapped_download"), {k: round(1e3 * v, 1) for k, v in d["e2e"]["breakdown_s"].items()})
PY
fi

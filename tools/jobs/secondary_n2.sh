mkdir -p gpurun_out/r02/scale
for c in fcos deeplab; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 2000)) bench.py --gpus 2 --config $c --steps 10 --warmup 3 > gpurun_out/r02/scale/bench_${c}_n2.json 2> gpurun_out/r02/scale/bench_${c}_n2.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02/scale/bench_${c}_n2.json").read().strip().splitlines()[-1]); print("$c", d["n_gpus"], d["value"], d["ms_per_step"], d["e2e"]["value"], d["config"]["collective"][:60])
except Exception as e: print("$c FAILED", e)
P
done

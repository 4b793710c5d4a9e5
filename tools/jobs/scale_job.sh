mkdir -p gpurun_out/r02/scale
run() {  # n, extra args..., tag
  n=$1; tag=$2; shift 2
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 2000)) bench.py --gpus $n "$@" > gpurun_out/r02/scale/$tag.json 2> gpurun_out/r02/scale/$tag.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02/scale/$tag.json").read().strip().splitlines()[-1]); print("$tag", d["n_gpus"], d["scaling"], d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("e2e_uint8_frames",{}).get("value"))
except Exception as e:
    print("$tag FAILED", e)
P
}
run 8 bench_n8 --steps 50 --warmup 5
run 8 bench_n8_strong --steps 50 --warmup 5 --scaling strong
run 4 bench_n4 --steps 50 --warmup 5
run 4 bench_n4_strong --steps 50 --warmup 5 --scaling strong
run 2 bench_n2 --steps 50 --warmup 5
run 2 bench_c3train_n2 --config c3train --steps 20 --warmup 5
run 8 bench_c3train_n8 --config c3train --steps 20 --warmup 5
tail -3 gpurun_out/r02/scale/*.err | cut -c1-300 | tail -30

mkdir -p gpurun_out/r02/knobs
b() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 30 --warmup 5 --cpu-steps 0 > gpurun_out/r02/knobs/$tag.json 2> gpurun_out/r02/knobs/$tag.err; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02/knobs/$tag.json").read().strip().splitlines()[-1]); print("$tag", d["value"], d["ms_per_step"], d["roofline"]["conv_ms_per_step"])
except Exception as e: print("$tag FAILED", e)
P
}
b base A=1
b wide256 CVB_BN64_WIDE_1X1=256
b wide128 CVB_BN64_WIDE_1X1=128
b base2 A=1
CVB_BN64_WIDE_1X1=256 CVB_PLAN_DEBUG=1 timeout 100 python tools/run_layer.py 256 256 1 1 40 40 64 3 2>&1 | tail -3 | cut -c1-400
timeout 100 python tools/run_layer.py 256 256 1 1 40 40 64 3 2>&1 | tail -1

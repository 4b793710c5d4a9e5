mkdir -p gpurun_out/r02/knobs
b() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --cpu-steps 0 > gpurun_out/r02/knobs/$tag.json 2> gpurun_out/r02/knobs/$tag.err; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02/knobs/$tag.json").read().strip().splitlines()[-1]); print("$tag", d["value"], d["ms_per_step"], d["roofline"]["conv_ms_per_step"])
except Exception as e: print("$tag FAILED", e)
P
}
b base A=1
b split128_0 CVB_SPLIT_N128=0
b bk_smalln_64 CVB_BK_SMALLN=64
b halo0 CVB_HALO=0
b maxchain_64 CVB_MAX_CHAIN=64
b pair0 CVB_MMA_PAIR=0
b base2 A=1

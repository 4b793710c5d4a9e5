mkdir -p gpurun_out/r02/knobs
b() { tag=$1; shift; env "$@" timeout 200 python bench.py --steps 40 --warmup 5 --cpu-steps 0 > gpurun_out/r02/knobs/$tag.json 2> gpurun_out/r02/knobs/$tag.err; python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02/knobs/$tag.json").read().strip().splitlines()[-1]); print("$tag", d["value"], d["ms_per_step"], d["roofline"]["conv_ms_per_step"], d["e2e"]["value"])
except Exception as e: print("$tag FAILED", e)
P
}
b pdl1 CVB_PDL=1
b base A=1
b pdl1b CVB_PDL=1

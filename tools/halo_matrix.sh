#!/usr/bin/env bash
# A/B matrix of the conv loader modes on single layers (run on the GPU box): prints one line per configuration.
set -u
cd "$(dirname "$0")/.."
run() {  # env..., then layer args
  local envs="$1"; shift
  local out
  out=$(env $envs CVB_PLAN_DEBUG=1 timeout 120 python tools/run_layer.py "$@" 5 2>&1)
  local plan=$(echo "$out" | grep "cvb plan" | head -1 | sed -E 's/.*(bn=[0-9]+ bk=[0-9]+ stages=[0-9]+ resident=[0-9]+ out_bufs=[0-9]+ ctas\/sm=[0-9]+).*(halo=.*)/\1 \2/')
  echo "$(echo "$out" | grep '^conv' ) | $envs | $plan"
}
for layer in "64 64 3 1 80 80 64" "32 32 3 1 160 160 64" "128 128 3 1 40 40 64" "32 64 3 2 320 320 64" "64 128 3 2 160 160 64" "256 256 3 1 20 20 64"; do
  run "CVB_HALO=0" $layer
  for mode in 2 1; do
    for bk in 64 32; do
      for ctas in 1 2; do
        for res in 1 0; do
          run "CVB_HALO=$mode CVB_HALO_BK=$bk CVB_HALO_CTAS=$ctas CVB_HALO_RES=$res" $layer
        done
      done
    done
  done
done

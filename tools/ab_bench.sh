#!/bin/bash
# A/B of library variants on one box: tools/ab_bench.sh "<variant.so>:<WN_LAYER_CHAIN>" ...   (developer aid)
cp lb_wavenet_b200/libwavenet_b200.so /tmp/keep.so
for spec in "$@"; do
  so=${spec%%:*}; mode=${spec##*:}
  cp lb_wavenet_b200/build/variants/$so.so lb_wavenet_b200/libwavenet_b200.so
  WN_LAYER_CHAIN=$mode timeout 120 python bench.py --steps 30 --warmup 3 --no-gen --no-cpu > gpurun_out/ab_${so}_$mode.json 2> gpurun_out/ab_${so}_$mode.err
  python - <<P
import json
try:
    d = json.load(open("gpurun_out/ab_${so}_$mode.json"))
    print("$so", "$mode", round(d["ms_per_step"], 3), round(d["value"] / 1e6, 2), {k: round(v["ms_per_step"], 3) for k, v in d["kernel_shares"].items()})
except Exception as e:
    print("$so $mode FAILED", e)
P
done
cp /tmp/keep.so lb_wavenet_b200/libwavenet_b200.so

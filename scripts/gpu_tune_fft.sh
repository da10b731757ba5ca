#!/bin/bash
# A/B builds of the transform-domain correlation kernel (threads per CTA, taps per block); each variant is a full library in build/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for lib in build/libhdn_b200_*.so; do
  tag=$(basename $lib .so | sed 's/libhdn_b200_//')
  HDN_B200_LIB=$PWD/$lib timeout 200 python bench.py --no-cpu --no-e2e --steps 10 > gpurun_out/tune_$tag.json 2> gpurun_out/tune_$tag.err
  HDN_B200_LIB=$PWD/$lib timeout 200 python bench.py --no-cpu --no-e2e --steps 10 --workload win15 > gpurun_out/tune_w15_$tag.json 2>> gpurun_out/tune_$tag.err
  python - "$tag" <<'PY'
import json, sys
t = sys.argv[1]
try:
    d = json.load(open("gpurun_out/tune_%s.json" % t)); w = json.load(open("gpurun_out/tune_w15_%s.json" % t))
    k = d["roofline"]["kernel_ms"]
    print("%-24s k1 %.3f  k2 %.3f  win15 %.3f" % (t, k["k1"], k["k2"], w["roofline"]["kernel_ms"]["k1"]))
except Exception as e:
    print(t, "FAILED", e)
PY
done

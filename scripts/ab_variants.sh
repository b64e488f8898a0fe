#!/bin/bash
set -x
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-residues --no-cufft --no-parity"
for v in "$@"; do
  if [ "$v" = default ]; then $B > gpurun_out/r02g_$v.json 2> gpurun_out/r02g_$v.err
  else FCD_B200_LIB=$PWD/build/ab/libfcd_$v.so $B > gpurun_out/r02g_$v.json 2> gpurun_out/r02g_$v.err; fi
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02g_$v.json").read().strip().splitlines()[-1])
    print("$v", round(d["value"],1), {k: round(v,2) for k,v in d["roofline"]["stage_us_per_frame"].items()})
except Exception as e:
    print("$v failed", e); print(open("gpurun_out/r02g_$v.err").read()[-1500:])
P
done

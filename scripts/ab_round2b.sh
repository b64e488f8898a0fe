#!/bin/bash
# Round-2 A/B on one B200: mask path after the exact constant division, and the twiddle-regeneration builds.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "mask" > gpurun_out/r02b_tmask.log 2>&1; tail -3 gpurun_out/r02b_tmask.log
python scripts/masked_workflow_bench.py > gpurun_out/r02b_masked.json 2> gpurun_out/r02b_masked.err; cat gpurun_out/r02b_masked.json
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-residues --no-cufft --no-parity"
$B > gpurun_out/r02b_ab_default.json 2> gpurun_out/r02b_ab_default.err
FCD_B200_LIB=$PWD/build/ab/libfcd_twregen1.so $B > gpurun_out/r02b_ab_regen1.json 2> gpurun_out/r02b_ab_regen1.err
FCD_B200_LIB=$PWD/build/ab/libfcd_twregen2.so $B > gpurun_out/r02b_ab_regen2.json 2> gpurun_out/r02b_ab_regen2.err
$B > gpurun_out/r02b_ab_default2.json 2> gpurun_out/r02b_ab_default2.err
for f in default regen1 regen2 default2; do python - <<P
import json
d=json.loads(open("gpurun_out/r02b_ab_$f.json").read().strip().splitlines()[-1])
print("$f", round(d["value"],1), {k: round(v,2) for k,v in d["roofline"]["stage_us_per_frame"].items()}, d["clocks"])
P
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02b_masked.csv python scripts/masked_workflow_bench.py > gpurun_out/r02b_ncu_masked.log 2>&1
tail -2 gpurun_out/r02b_ncu_masked.log

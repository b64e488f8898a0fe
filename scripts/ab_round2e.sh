#!/bin/bash
set -x
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-residues --no-cufft --no-parity --unwrap off"
$B > gpurun_out/r02e_default.json 2> gpurun_out/r02e_default.err
for i in 1 2 3 4 5; do FCD_B200_LIB=$PWD/build/ab/libfcd_probe$i.so $B > gpurun_out/r02e_probe$i.json 2> gpurun_out/r02e_probe$i.err; done
for f in default probe1 probe2 probe3 probe4 probe5; do python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02e_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), {k: round(v,2) for k,v in d["roofline"]["stage_us_per_frame"].items()})
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r02e_$f.err").read()[-1500:])
P
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02e_masked.csv python scripts/masked_workflow_bench.py > gpurun_out/r02e_ncu_masked.log 2>&1

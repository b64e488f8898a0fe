#!/bin/bash
set -x
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-residues --no-cufft --no-parity"
$B > gpurun_out/r02f_default.json 2> gpurun_out/r02f_default.err
FCD_B200_LIB=$PWD/build/ab/libfcd_fg.so $B > gpurun_out/r02f_fg.json 2> gpurun_out/r02f_fg.err
for f in default fg; do python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02f_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), {k: round(v,2) for k,v in d["roofline"]["stage_us_per_frame"].items()})
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r02f_$f.err").read()[-1500:])
P
done
python scripts/variant_checksum.py 2048 > gpurun_out/r02f_sum_default.json 2> gpurun_out/r02f_sum_default.err
FCD_B200_LIB=$PWD/build/ab/libfcd_fg.so python scripts/variant_checksum.py 2048 > gpurun_out/r02f_sum_fg.json 2> gpurun_out/r02f_sum_fg.err
cat gpurun_out/r02f_sum_default.json gpurun_out/r02f_sum_fg.json; tail -n 3 gpurun_out/r02f_sum_fg.err

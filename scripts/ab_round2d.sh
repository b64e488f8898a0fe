#!/bin/bash
set -x
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-residues --no-cufft --no-parity"
$B > gpurun_out/r02d_ab_default.json 2> gpurun_out/r02d_ab_default.err
FCD_B200_LIB=$PWD/build/ab/libfcd_seq6.so $B > gpurun_out/r02d_ab_seq6.json 2> gpurun_out/r02d_ab_seq6.err
FCD_B200_LIB=$PWD/build/ab/libfcd_seq5.so $B > gpurun_out/r02d_ab_seq5.json 2> gpurun_out/r02d_ab_seq5.err
for f in default seq6 seq5; do python - <<P
import json
try:
    d=json.loads(open("gpurun_out/r02d_ab_$f.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"],1), {k: round(v,2) for k,v in d["roofline"]["stage_us_per_frame"].items()}, d["clocks"])
except Exception as e:
    print("$f failed", e); print(open("gpurun_out/r02d_ab_$f.err").read()[-1500:])
P
done
python scripts/variant_checksum.py 2048 > gpurun_out/r02d_sum_default.json 2> gpurun_out/r02d_sum_default.err
FCD_B200_LIB=$PWD/build/ab/libfcd_seq6.so python scripts/variant_checksum.py 2048 > gpurun_out/r02d_sum_seq6.json 2> gpurun_out/r02d_sum_seq6.err
FCD_B200_LIB=$PWD/build/ab/libfcd_seq5.so python scripts/variant_checksum.py 2048 > gpurun_out/r02d_sum_seq5.json 2> gpurun_out/r02d_sum_seq5.err
FCD_B200_LIB=$PWD/build/ab/libfcd_seq6.so python scripts/variant_checksum.py 1024 > gpurun_out/r02d_sum_seq6_1024.json 2> gpurun_out/r02d_sum_seq6_1024.err
python scripts/variant_checksum.py 1024 > gpurun_out/r02d_sum_default_1024.json 2> gpurun_out/r02d_sum_default_1024.err
cat gpurun_out/r02d_sum_*.json; tail -3 gpurun_out/r02d_sum_*.err

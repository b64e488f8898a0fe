#!/bin/bash
# A/B of two builds, then the record of the faster one: bench line (parity on the timed frames), output digests, GPU tests.
mkdir -p gpurun_out
B="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-residues --no-cufft --no-parity"
for v in A B; do FCD_B200_LIB=$PWD/build/ab/libfcd_$v.so $B > gpurun_out/r02i_$v.json 2> gpurun_out/r02i_$v.err; done
W=$(python - <<'P'
import json
r={}
for v in "AB":
    d=json.loads(open(f"gpurun_out/r02i_{v}.json").read().strip().splitlines()[-1])
    r[v]=d["value"]; print(v, round(d["value"],1), {k: round(x,2) for k,x in d["roofline"]["stage_us_per_frame"].items()}, file=__import__("sys").stderr)
print(max(r, key=r.get))
P
)
echo "winner $W" | tee gpurun_out/r02i_winner.txt
export FCD_B200_LIB=$PWD/build/ab/libfcd_$W.so
timeout 100 python bench.py --no-cufft --no-residues > gpurun_out/r02i_final.json 2> gpurun_out/r02i_final.err; tail -c 400 gpurun_out/r02i_final.json
timeout 40 python scripts/variant_checksum.py 2048 > gpurun_out/r02i_sum.json 2>/dev/null; cat gpurun_out/r02i_sum.json
timeout 120 python -m pytest tests -m gpu -q -x > gpurun_out/r02i_tests.log 2>&1; tail -2 gpurun_out/r02i_tests.log

#!/bin/bash
# End-of-session record on one B200: GPU test suite, the contract bench line, ncu launch list and one full capture.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02b_final_tests.log 2>&1; tail -2 gpurun_out/r02b_final_tests.log
python bench.py > gpurun_out/r02b_final.json 2> gpurun_out/r02b_final.err; rc=$?; tail -c 600 gpurun_out/r02b_final.json
if [ $rc -eq 0 ]; then
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r02b.csv \
    python bench.py --steps 2 --warmup 3 --frames 256 --no-e2e --no-cpu-baseline --no-cufft --no-parity --no-residues > gpurun_out/r02b_ncu_launches.log 2>&1
  timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"RowFwd|ColBand|RowDemod|RowLink|ColIntegrate|RowInv" -s 12 -c 6 -f -o gpurun_out/r02b \
    python scripts/profile_run.py 2048 32 > gpurun_out/r02b_ncu_full.log 2>&1
  tail -2 gpurun_out/r02b_ncu_full.log
fi
ls -l gpurun_out | tail -8

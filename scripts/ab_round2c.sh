#!/bin/bash
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x > gpurun_out/r02c_tests.log 2>&1; tail -3 gpurun_out/r02c_tests.log
python scripts/masked_workflow_bench.py > gpurun_out/r02c_masked.json 2> gpurun_out/r02c_masked.err; cat gpurun_out/r02c_masked.json
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"BoxRowsWarp|BoxLines|LabelInit|LabelMerge" -c 6 -o gpurun_out/r02c_mask python scripts/masked_workflow_bench.py > gpurun_out/r02c_ncu_mask.log 2>&1
tail -2 gpurun_out/r02c_ncu_mask.log
ls -l gpurun_out/

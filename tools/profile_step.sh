#!/usr/bin/env bash
# ncu evidence for ONE training step of the bench command (run on the GPU box through gpurun, one GPU):
#   1. launch list  (gpu__time_duration.sum, --clock-control none)     -> gpurun_out/<tag>_launches.csv
#   2. --set full capture of the same step + raw-page CSV               -> gpurun_out/<tag>_full.ncu-rep / _full_raw.csv
#   3. ordered C-ABI call log of the profiled step (bench.py --call-log) -> gpurun_out/<tag>_calls.json
# tools/ncu_join.py joins 2 + 3 into profiles/<tag>_ncu_calls.json (what bench.py reads `roofline.traffic` from).
# Numbers printed by a run under ncu are never bench values.
set -u
TAG=${1:-round2}
OUT=gpurun_out
mkdir -p $OUT
NCU=${NCU:-ncu}
$NCU --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
     --log-file $OUT/${TAG}_launches.csv python bench.py --ncu-steps 1 --warmup 3 > $OUT/${TAG}_launches.log 2>&1
echo "launch list rc=$?"
$NCU --set full --clock-control none --import-source on --profile-from-start off -o $OUT/${TAG}_full -f \
     python bench.py --ncu-steps 1 --warmup 3 --call-log $OUT/${TAG}_calls.json > $OUT/${TAG}_full.log 2>&1
echo "full capture rc=$?"
$NCU -i $OUT/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_full_raw.csv 2> $OUT/${TAG}_full_raw.err
echo "raw export rc=$?"
python tools/ncu_join.py $OUT/${TAG}_full_raw.csv $OUT/${TAG}_calls.json $OUT/${TAG}_ncu_calls.json
echo "join rc=$?"

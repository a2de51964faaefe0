#!/usr/bin/env bash
# ncu evidence for ONE training step of the bench command (run on the GPU box through gpurun, one GPU):
#   1. launch list (gpu__time_duration.sum, --clock-control none)            -> gpurun_out/<tag>_launches.csv
#   2. DRAM / pipe counters of EVERY kernel of the step (a few passes each)   -> gpurun_out/<tag>_step_raw.csv
#      + the ordered C-ABI call log of the same step (bench.py --call-log)   -> gpurun_out/<tag>_calls.json
#      joined by tools/ncu_join.py                                            -> gpurun_out/<tag>_ncu_calls.json
#   3. `--set full` capture of the dominant kernels (name filter, first launches of the step)
#                                                                             -> gpurun_out/<tag>_full_raw.csv (+ .ncu-rep if small)
# Numbers printed by a run under ncu are never bench values.
set -u
TAG=${1:-round2}
FULL_FILTER=${2:-'regex:tc_gemm|gat_bwd_edge|stream_kernel|proj_|pool_maxmean|gat_alpha'}
FULL_COUNT=${3:-16}
OUT=gpurun_out
mkdir -p $OUT
NCU=${NCU:-ncu}
METRICS=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,sm__inst_issued.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__inst_executed.sum
$NCU --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
     --log-file $OUT/${TAG}_launches.csv python bench.py --ncu-steps 1 --warmup 3 > $OUT/${TAG}_launches.log 2>&1
echo "launch list rc=$?"
python tools/summarize_launches.py $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches.txt 2>&1
$NCU --metrics $METRICS --clock-control none --profile-from-start off -o $OUT/${TAG}_step -f \
     python bench.py --ncu-steps 1 --warmup 3 --call-log $OUT/${TAG}_calls.json > $OUT/${TAG}_step.log 2>&1
echo "step counters rc=$?"
$NCU -i $OUT/${TAG}_step.ncu-rep --page raw --csv > $OUT/${TAG}_step_raw.csv 2> $OUT/${TAG}_step_raw.err
python tools/ncu_join.py $OUT/${TAG}_step_raw.csv $OUT/${TAG}_calls.json $OUT/${TAG}_ncu_calls.json
echo "join rc=$?"
rm -f $OUT/${TAG}_step.ncu-rep
if [ "$FULL_COUNT" != "0" ]; then
  $NCU --set full --clock-control none --profile-from-start off -k "$FULL_FILTER" -c $FULL_COUNT -o $OUT/${TAG}_full -f \
       python bench.py --ncu-steps 1 --warmup 3 > $OUT/${TAG}_full.log 2>&1
  echo "full capture rc=$?"
  $NCU -i $OUT/${TAG}_full.ncu-rep --page raw --csv > $OUT/${TAG}_full_raw.csv 2> $OUT/${TAG}_full_raw.err
  SZ=$(stat -c %s $OUT/${TAG}_full.ncu-rep 2>/dev/null || echo 0)
  if [ "$SZ" -gt 25000000 ]; then rm -f $OUT/${TAG}_full.ncu-rep; echo "full .ncu-rep dropped ($SZ bytes), raw CSV kept"; fi
fi
du -sh $OUT

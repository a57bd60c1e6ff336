#!/bin/bash
# Round profile on the GPU box (single GPU; every ncu pass only after the plain run of the same command exited 0):
#   1. plain run of the exact command                       -> gpurun_out/prof_plain.json
#   2. launch list: ncu --metrics gpu__time_duration.sum     -> gpurun_out/launches.csv
#   3. one `ncu --set full` capture per hot kernel           -> gpurun_out/full_<kernel>.ncu-rep
# Summaries are made afterwards with tools/ncu_launches.py and tools/ncu_summary.py and copied to profiles/.
set -u
CMD="python bench.py --steps 1 --warmup 1 --gops 4 --no-cpu-baseline --no-extra --no-verify --no-transcode"
$CMD > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err || { echo "plain run failed"; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
export VCPENC_STREAMS=1
FULL="python bench.py --steps 1 --warmup 1 --gops 8 --no-cpu-baseline --no-extra --no-verify --no-transcode"
$FULL > gpurun_out/prof_plain8.json 2> gpurun_out/prof_plain8.err || { echo "plain run (8 GOPs) failed"; exit 1; }
for k in me_prepass_kernel me_prepass_l0_kernel me_refine p_recon hpel_planes deblock_kernel cavlc_mb; do
  skip=5; case $k in me_prepass*) skip=1;; esac
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip $skip -c 1 -f -o gpurun_out/full_$k \
     $FULL > gpurun_out/ncu_full_$k.log 2>&1
  echo "$k rc=$?"
done
$FULL --entropy 1 --t8x8 1 > gpurun_out/prof_plain8c.json 2> gpurun_out/prof_plain8c.err || { echo "plain CABAC run failed"; exit 1; }
for k in cabac_bins cabac_gather cabac_encode; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 2 -c 1 -f -o gpurun_out/full_$k \
     $FULL --entropy 1 --t8x8 1 > gpurun_out/ncu_full_$k.log 2>&1
  echo "$k rc=$?"
done

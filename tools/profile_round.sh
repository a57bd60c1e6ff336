#!/bin/bash
# Round profile on the GPU box (single GPU; each ncu pass only after the plain run exited 0):
#   1. plain run of the exact command                       -> gpurun_out/prof_plain.json
#   2. launch list: ncu --metrics gpu__time_duration.sum     -> gpurun_out/launches.csv
#   3. one `ncu --set full` capture per hot kernel           -> gpurun_out/full_<kernel>.ncu-rep
# Summaries are made afterwards with tools/ncu_launches.py and tools/ncu_summary.py and copied to profiles/.
set -u
CMD="python bench.py --steps 1 --warmup 1 --gops 4 --no-cpu-baseline"
$CMD > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err || { echo "plain run failed"; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
export VCPENC_STREAMS=1
for k in me_prepass me_refine p_recon hpel_planes deblock_kernel cavlc_mb i_recon k1_yuv420p; do
  skip=5; case $k in me_prepass|k1_yuv420p) skip=1;; i_recon) skip=2;; esac
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip $skip -c 1 -f -o gpurun_out/full_$k \
     python bench.py --steps 1 --warmup 1 --gops 8 --no-cpu-baseline > gpurun_out/ncu_full_$k.log 2>&1
  echo "$k rc=$?"
done
for k in cabac_bins cabac_encode; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 2 -c 1 -f -o gpurun_out/full_$k \
     python bench.py --steps 1 --warmup 1 --gops 8 --no-cpu-baseline --entropy 1 > gpurun_out/ncu_full_$k.log 2>&1
  echo "$k rc=$?"
done

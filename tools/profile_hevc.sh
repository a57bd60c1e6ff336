#!/bin/bash
# HEVC profile on the GPU box (single GPU; ncu passes only after the plain run exited 0):
#   plain run -> gpurun_out/hevc_plain.json; launch list -> gpurun_out/hevc_launches.csv;
#   `ncu --set full` of hevc_p_recon / hevc_bins / hevc_i_recon -> gpurun_out/full_<kernel>.ncu-rep
set -u
CMD="python bench.py --codec hevc --steps 1 --warmup 1 --gops 4 --no-cpu-baseline"
$CMD > gpurun_out/hevc_plain.json 2> gpurun_out/hevc_plain.err || { echo "plain run failed"; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/hevc_launches.csv $CMD > gpurun_out/ncu_hevc_launches.log 2>&1
echo "launch list rc=$?"
export VCPENC_STREAMS=1
for k in hevc_p_recon hevc_bins hevc_i_recon; do
  skip=5; [ "$k" = "hevc_i_recon" ] && skip=2
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip $skip -c 1 -f -o gpurun_out/full_$k \
     python bench.py --codec hevc --steps 1 --warmup 1 --gops 8 --no-cpu-baseline > gpurun_out/ncu_full_$k.log 2>&1
  echo "$k rc=$?"
done

#!/bin/bash
# One `ncu --set full` capture per hot kernel (single GPU, never under torchrun).  Usage on the
# GPU box: bash tools/ncu_capture.sh <tag> [kernel ...]; reports land in gpurun_out/prof_<tag>_<kernel>.ncu-rep
tag=$1; shift
kernels=${@:-"me_refine me_prepass p_recon hpel_planes"}
export VCPENC_STREAMS=1
for k in $kernels; do
  skip=5; [ "$k" = "me_prepass" ] && skip=1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip $skip -c 1 \
    -f -o gpurun_out/prof_${tag}_$k python bench.py --steps 1 --warmup 1 --gops 8 --no-cpu-baseline > gpurun_out/ncu_${tag}_$k.log 2>&1
  echo "$k rc=$?"
done

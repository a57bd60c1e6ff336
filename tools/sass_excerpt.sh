#!/bin/bash
# SASS evidence for the TMA-staged search windows: tools/sass_excerpt.sh > profiles/r02_sass_tma.txt
LIB=video_codec_pipeline_b200/lib/libvcpenc.so
echo "# cuobjdump -sass $LIB -- TMA / mbarrier instructions of the motion-search kernels"
echo "# PTX -> SASS: cp.async.bulk.tensor -> UTMALDG, mbarrier.init -> SYNCS.EXCH.64, mbarrier.arrive.expect_tx -> SYNCS.ARRIVE.TRANS64,"
echo "#              mbarrier.try_wait.parity -> SYNCS.PHASECHK.TRANS64.TRYWAIT, the SAD work -> VABSDIFF4.U8.ACC"
for k in me_prepass_kernel me_refine_kernel; do
  echo; echo "## $k"
  cuobjdump -sass $LIB | awk -v k=$k '/Function :/{f=($0 ~ k)} f' | sed -E 's/\s+\/\* 0x[0-9a-f]+ \*\///' > /tmp/_k.sass
  grep -E "UTMALDG|SYNCS.ARRIVE|SYNCS.PHASECHK" /tmp/_k.sass | sed -E 's/^\s+//' | cut -c1-140
  echo "# mnemonic counts:"
  for m in UTMALDG SYNCS.EXCH SYNCS.ARRIVE SYNCS.PHASECHK VABSDIFF4 LDS LDG STS REDUX SHFL; do printf "#   %-16s %s\n" $m $(grep -c "$m" /tmp/_k.sass); done
  printf "#   %-16s %s\n" "instructions" $(grep -cE "^\s+/\*[0-9a-f]{4}\*/" /tmp/_k.sass)
done
echo; echo "## whole library"
for m in UTMALDG UTMASTG UBLKCP LDGSTS SYNCS VABSDIFF4 HMMA UTCHMMA; do printf "#   %-10s %s\n" $m $(cuobjdump -sass $LIB | grep -c "$m"); done

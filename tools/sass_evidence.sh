#!/bin/bash
# Counts the Blackwell/Hopper-class instructions per kernel in the built objects (TMA tile loads UTMALDG, bulk copies UBLKCP,
# cp.async LDGSTS, mbarrier SYNCS, cluster barriers UCGABAR) -> profiles/r02_sass_evidence.md
for f in orb ba_sweep ba_pcg_bcsr ba bow match; do
  echo "## gslam_b200/build/$f.o"
  cuobjdump -sass gslam_b200/build/$f.o | grep -E "Function :|UTMALDG|UBLKCP|LDGSTS|SYNCS|UCGABAR|POPC"
done

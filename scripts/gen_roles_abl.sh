#!/bin/bash
# developer aid: per-role clocks of CTA 0 under work-skipping switches (library built with -DZ3D_GEN_TIMING -DGT_BLOCK=0 -DZ3D_GEN_DEBUG -DZ3D_GEN_ABL)
cp zernike3d_b200/libz3d_b200.so /tmp/lib_keep.so
cp zernike3d_b200/libz3d_b200_dbg.so zernike3d_b200/libz3d_b200.so
for d in ${2:-0 61441}; do
  echo "== DBG $d"; Z3D_GEN_DBG=$d timeout 120 python scripts/gen_roles.py ${1:-128 50 64}
done
cp /tmp/lib_keep.so zernike3d_b200/libz3d_b200.so

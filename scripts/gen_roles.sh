#!/bin/bash
# developer aid: per-role clocks of chosen CTAs (library built with -DZ3D_GEN_TIMING as libz3d_b200_dbg.so); $1 = "dim order B", $2 = CTA list
cp zernike3d_b200/libz3d_b200.so /tmp/lib_keep.so
cp zernike3d_b200/libz3d_b200_dbg.so zernike3d_b200/libz3d_b200.so
for d in ${2:-0}; do
  echo "== CTA $d"; Z3D_GEN_DBG=$d timeout 120 python scripts/gen_roles.py ${1:-128 50 128}
done
cp /tmp/lib_keep.so zernike3d_b200/libz3d_b200.so

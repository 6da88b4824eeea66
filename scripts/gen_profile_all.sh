#!/bin/bash
# developer aid: per-role clocks (libz3d_b200_tim.so = -DZ3D_GEN_TIMING) and work-skipping ablations (libz3d_b200_abl.so =
# -DZ3D_GEN_DEBUG -DZ3D_GEN_ABL) of the table-free kernel in one GPU-box pass; $1 = "dim order B"
cp zernike3d_b200/libz3d_b200.so /tmp/lib_keep.so
cp zernike3d_b200/libz3d_b200_tim.so zernike3d_b200/libz3d_b200.so
for d in ${2:-0 1 74 147}; do
  echo "== CTA $d"; Z3D_GEN_DBG=$d timeout 120 python scripts/gen_roles.py ${1:-128 50 64}
done
cp zernike3d_b200/libz3d_b200_abl.so zernike3d_b200/libz3d_b200.so
for d in 0 1 2048 4096 8192 16384 32768 4097 8193 16385 12289 28673 61441; do
  Z3D_GEN_DBG=$d timeout 120 python scripts/gen_time.py ${1:-128 50 64} | sed "s/^/DBG $d: /"
done
cp /tmp/lib_keep.so zernike3d_b200/libz3d_b200.so

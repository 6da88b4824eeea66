#!/bin/bash
# developer aid: timing ablations of the table-free kernel (library built with -DZ3D_GEN_DEBUG as libz3d_b200_dbg.so)
# bits: 1 no read-back, 2048 one MMA per segment, 4096 no a-panel work, 8192 no T work, 16384 no MMAs, 32768 no seed work
cp zernike3d_b200/libz3d_b200.so /tmp/lib_keep.so
cp zernike3d_b200/libz3d_b200_dbg.so zernike3d_b200/libz3d_b200.so
for d in ${2:-0 1 4096 8192 16384 4097 8193 16385 12289 28673 61441}; do
  Z3D_GEN_DBG=$d timeout 120 python scripts/gen_time.py ${1:-128 50 128} | sed "s/^/DBG $d: /"
done
cp /tmp/lib_keep.so zernike3d_b200/libz3d_b200.so

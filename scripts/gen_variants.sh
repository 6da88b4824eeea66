#!/bin/bash
# developer aid: time library variants zernike3d_b200/libz3d_b200_<name>.so ($2..) on the shape $1 = "dim order B"
cp zernike3d_b200/libz3d_b200.so /tmp/lib_keep.so
shape=$1; shift
for v in "$@"; do
  cp zernike3d_b200/libz3d_b200_$v.so zernike3d_b200/libz3d_b200.so
  timeout 120 python scripts/gen_time.py $shape | sed "s/^/variant $v: /"
done
cp /tmp/lib_keep.so zernike3d_b200/libz3d_b200.so

#!/bin/bash
# developer aid: correctness + time of library variants zernike3d_b200/libz3d_b200_<name>.so ($2..) on the cases $1 = "dim,order,B ..."
cp zernike3d_b200/libz3d_b200.so /tmp/lib_keep.so
cases=$1; shift
for v in "$@"; do
  cp zernike3d_b200/libz3d_b200_$v.so zernike3d_b200/libz3d_b200.so
  timeout 200 python scripts/gen2_check.py $cases | sed "s/^/variant $v: /"
done
cp /tmp/lib_keep.so zernike3d_b200/libz3d_b200.so

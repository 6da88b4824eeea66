#!/bin/bash
# developer aid: correctness + time of library variants zernike3d_b200/libz3d_b200_<name>.so ($2.., "cur" = the product build) on the cases $1
cp zernike3d_b200/libz3d_b200.so /tmp/lib_keep.so
cases=$1; shift
for v in "$@"; do
  if [ "$v" != cur ]; then cp zernike3d_b200/libz3d_b200_$v.so zernike3d_b200/libz3d_b200.so; else cp /tmp/lib_keep.so zernike3d_b200/libz3d_b200.so; fi
  timeout 300 python scripts/gen2_check.py $cases 2>&1 | sed "s/^/variant $v: /"
done
cp /tmp/lib_keep.so zernike3d_b200/libz3d_b200.so

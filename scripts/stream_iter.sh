#!/bin/bash
# developer loop on the GPU box: parity check, timing, one ncu capture of the streamed-T kernel
tag=${1:-x}
timeout 300 python scripts/batched_check.py d32_o20 c1_d64_o50:64 > gpurun_out/s_$tag.log 2>&1; echo rc=$? >> gpurun_out/s_$tag.log
timeout 200 python scripts/time_batched.py 128 50 64 >> gpurun_out/s_$tag.log 2>&1; echo rc=$? >> gpurun_out/s_$tag.log
if [ "$2" = "ncu" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_decompose_stream -s 1 -c 1 -o gpurun_out/r1_stream_$tag -f python scripts/prof_run.py 128 50 64 2 dec > gpurun_out/ncu_s.log 2>&1
fi
tail -6 gpurun_out/s_$tag.log

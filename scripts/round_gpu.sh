#!/bin/bash
# one GPU-box pass: GPU tests, smoke, bench, launch list, full ncu captures of the dominant kernels; $1 = "ncu" adds the profiler passes
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/r2e_bench_1gpu.json 2> gpurun_out/bench1.err; echo "bench rc=$?" >> gpurun_out/bench1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2e_bench_reference_arm.json 2> gpurun_out/benchref.err
if [ "$1" = "ncu" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2e_launches_bench.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_decompose_gen -s 1 -c 1 -o gpurun_out/r2e_gen_full -f python scripts/prof_run.py 128 50 64 2 dec > gpurun_out/ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_synth_gen -s 8 -c 1 -o gpurun_out/r2e_synth_full -f python scripts/synth_check.py 128,50,64 > gpurun_out/ncu_full_synth.log 2>&1
fi
tail -3 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/smoke.log; cut -c1-400 gpurun_out/r2e_bench_1gpu.json; tail -2 gpurun_out/bench1.err

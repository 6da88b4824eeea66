"""Summarise an .ncu-rep here (no GPU): key raw metrics + the SASS instructions with the most stall samples.
usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep [top_n]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
h, u, v = r[0], r[1], r[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_tensor.sum", "smsp__cycles_active.avg"]
for i, k in enumerate(h):
    if k in want:
        print(f"{k:90s} {u[i]:10s} {v[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h, data = rows[1], rows[2:]
ia, isamp, iex = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
stall = [i for i, k in enumerate(h) if k.startswith("stall_") and "Not Issued" not in k]
tot = sum(int(x[isamp]) for x in data)
print("total samples", tot, "instructions", sum(int(x[iex]) for x in data))
agg = {}
for x in data:
    for i in stall:
        agg[h[i]] = agg.get(h[i], 0) + int(x[i])
print("stall totals:", ", ".join(f"{k[6:]} {100 * c / tot:.1f}%" for k, c in sorted(agg.items(), key=lambda kv: -kv[1])[:10]))
for n, x in sorted(enumerate(data), key=lambda kv: -int(kv[1][isamp]))[:topn]:
    st = sorted(((int(x[i]), h[i][6:]) for i in stall), reverse=True)[:2]
    print(f"{n:5d} {int(x[isamp]) * 100 / tot:5.1f}% ex={int(x[iex]):>10} {x[ia].strip()[:72]:72s} {st}")

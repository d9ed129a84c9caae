"""Summary of one ncu report (--set full --import-source on): key metrics + per-instruction warp-stall samples.
   python tools/ncu_stalls.py report.ncu-rep [top_n]"""
import csv, collections, subprocess, sys, io
rep = sys.argv[1]; top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(raw)))
hdr, units, row = r[0], r[1], r[2]
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "sm__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "launch__registers_per_thread", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.per_cycle_active", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
print("kernel:", row[hdr.index("Kernel Name")])
for w in want:
    if w in hdr:
        i = hdr.index(w); print(f"{w} [{units[i]}] = {row[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(io.StringIO(src)))
hi = next(i for i, x in enumerate(r) if x and x[0] == "Address")
hdr, rows = r[hi], [x for x in r[hi + 1:] if len(x) == len(r[hi])]
si, ex = hdr.index("# Samples"), hdr.index("Instructions Executed")
cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
for x in rows:
    for i in cols:
        tot[hdr[i][6:]] += int(x[i] or 0)
print("warp samples", sum(int(x[si] or 0) for x in rows), "; by stall reason:", ", ".join(f"{k} {v}" for k, v in tot.most_common(10)))
print("top instructions by samples (samples, executions, SASS, dominant stalls):")
for x in sorted(rows, key=lambda x: -int(x[si] or 0))[:top_n]:
    st = sorted(((int(x[i] or 0), hdr[i][6:]) for i in cols if x[i] not in ("", "0")), reverse=True)[:3]
    print(f"  {x[si]:>5} {x[ex]:>8}  {x[1].strip()[:64]:<64} {st}")

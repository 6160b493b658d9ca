"""Summarise an .ncu-rep (raw page + source page): key metrics, stall mix, instruction mix, hottest lines."""
import collections, csv, io, re, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get("Kernel Name", "")[:90])
    for k in ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
              "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
              "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size",
              "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
              "sm__inst_executed.avg.per_cycle_elapsed", "smsp__inst_executed.sum",
              "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
              "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
              "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum"):
        if k in d: print(f"  {k:70s} {d[k]} {units[hdr.index(k)]}")
    st = {h: float(d[h].replace(",", "")) for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and d[h]}
    print("  stalls:", ", ".join(f"{k.split('issue_stalled_')[1].split('_per_issue')[0]}={v:.2f}" for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]; ix = {k: i for i, k in enumerate(h)}
byop, tot, lines = collections.Counter(), 0, []
for r in rows[2:]:
    if len(r) < 10 or not r[ix["Instructions Executed"]].isdigit(): continue
    n, s = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
    byop[m.group(2) if m else "?"] += n; tot += n; lines.append((s, n, r[ix["Source"]].strip()))
print("instruction mix:", ", ".join(f"{k} {100*v/tot:.1f}%" for k, v in byop.most_common(18)))
lines.sort(reverse=True)
print("hottest (samples, executed, sass):")
for s, n, t in lines[:int(sys.argv[2]) if len(sys.argv) > 2 else 14]: print("  ", s, n, t)

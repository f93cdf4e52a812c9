"""Summarise an .ncu-rep (read here, no GPU needed) into a small text file for profiles/.
usage: ncu_summary.py report.ncu-rep out.txt"""
import collections
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__cycles_active.avg", "l1tex__t_set_accesses_pipe_lsu_mem_global_op_atom.sum", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
        "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_requests_srcunit_tex_op_atom.sum",
        "lts__t_requests_srcunit_tex_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "sm__cycles_elapsed.max", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed"]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
lines = []
if len(rows) > 2:
    hdr, units = rows[0], rows[1]
    only = sys.argv[3] if len(sys.argv) > 3 else ""
    for r in rows[2:]:
        if only and only not in dict(zip(hdr, r)).get("Kernel Name", ""):
            continue
        d = dict(zip(hdr, r))
        lines.append("kernel: %s" % d.get("Kernel Name"))
        for k in KEEP:
            if k in d:
                lines.append("  %-66s %s %s" % (k, d[k], units[hdr.index(k)]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur, ixm, n_hdr = None, None, 0
agg, text = collections.defaultdict(lambda: [0, 0]), {}
stall = collections.Counter()
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        ixm = {h: i for i, h in enumerate(r)}
        n_hdr = len(r)
        continue
    if ixm is None or len(r) != n_hdr:
        continue
    try:
        line = int(r[0])
        ie, ns = int(r[ixm["Instructions Executed"]] or 0), int(r[ixm["# Samples"]] or 0)
    except ValueError:
        continue
    a = agg[(cur, line)]
    a[0] += ie
    a[1] += ns
    text[(cur, line)] = r[1].strip()
    for h, i in ixm.items():
        if h.startswith("stall_") and "Not Issued" not in h:
            try:
                stall[h] += int(r[i] or 0)
            except ValueError:
                pass
tot = sum(v[1] for v in agg.values()) or 1
lines.append("")
lines.append("warp stall sampling, all warps (%d samples):" % tot)
for h, v in stall.most_common(10):
    lines.append("  %-28s %6.1f %%" % (h, 100.0 * v / tot))
lines.append("")
lines.append("instructions executed (warp level): %d" % sum(v[0] for v in agg.values()))
lines.append("top source lines by samples:")
for (f, l), v in sorted(agg.items(), key=lambda x: -x[1][1])[:40]:
    lines.append("  %5.1f %%  instr %9d  %s:%d  %s" % (100.0 * v[1] / tot, v[0], f, l, text.get((f, l), "")[:110]))
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:30]))

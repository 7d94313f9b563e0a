#!/usr/bin/env python
"""Summarise an .ncu-rep (read with `ncu -i`) into a small text file for profiles/: key SM / tensor / DRAM metrics, the
warp-stall breakdown and the hottest SASS instructions of EVERY profiled launch.   usage: ncu_summary.py rep out.txt"""
import csv
import io
import subprocess
import sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.avg", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors.sum",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Address":
        cur = {"h": r, "rows": []}
        blocks.append(cur)
    elif cur is not None and len(r) > 5:
        cur["rows"].append(r)
with open(out, "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on   (%s)\n" % rep.split("/")[-1])
    for li, vals in enumerate(rows[2:]):
        d = dict(zip(hdr, vals))
        f.write("\n===== launch %d =====\n" % li)
        for k in keys:
            if k in d:
                f.write("%-72s %s %s\n" % (k, d[k], units[hdr.index(k)]))
        st = [(k.replace("smsp__pcsamp_warps_issue_stalled_", ""), float(d[k].replace(",", ""))) for k in hdr
              if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k and d[k] not in ("", "n/a")]
        st.sort(key=lambda x: -x[1])
        tot = sum(v for _, v in st) or 1
        f.write("\n# warp stall samples (all)\n")
        for k, v in st[:10]:
            f.write("%-32s %8.0f %5.1f%%\n" % (k, v, 100 * v / tot))
        if li < len(blocks):
            h = blocks[li]["h"]
            so, ws, ie = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
            data = []
            for r in blocks[li]["rows"]:
                try:
                    data.append((int(r[ws] or 0), r[so], r[ie]))
                except Exception:
                    pass
            t = sum(x[0] for x in data) or 1
            f.write("\n# hottest SASS instructions (samples, %, instruction, executions)\n")
            for i in sorted(range(len(data)), key=lambda i: -data[i][0])[:12]:
                prev = data[i - 1][1][:48] if i else ""
                f.write("%6d %5.1f%%  %-64s exec=%-9s | prev: %s\n" % (data[i][0], 100 * data[i][0] / t, data[i][1][:64], data[i][2], prev))

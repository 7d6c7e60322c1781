#!/usr/bin/env python3
"""Turn an .ncu-rep (ncu --set full --import-source on) into the two text files committed next to it:
    <name>.summary.csv   selected raw metrics of the launch
    <name>.hotlines.txt  warp-stall samples by source line (top 40) and by stall reason
usage: python profiles/summarize.py gpurun_out/<file>.ncu-rep profiles/<name> [object-file-with-lineinfo.o]
The optional object file (same build) supplies the address -> source line map via nvdisasm -g."""
import collections
import csv
import io
import os
import re
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "launch__block_size", "launch__grid_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__cycles_active.avg", "sm__cycles_elapsed.max"]


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2]


def main():
    rep, name = sys.argv[1], sys.argv[2]
    obj = sys.argv[3] if len(sys.argv) > 3 else None
    hdr, units, vals = raw(rep)
    with open(name + ".summary.csv", "w") as fh:
        fh.write("metric,unit,value\n")
        for h, u, v in sorted(zip(hdr, units, vals)):
            if h in KEEP or (h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")):
                fh.write("%s,%s,%s\n" % (h, u, v))
        for h, v in zip(hdr, vals):
            if h == "Kernel Name":
                fh.write("kernel,,%s\n" % v.replace(",", ";"))
    # stall samples by line
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        return
    h = rows[1]
    ix = {k: i for i, k in enumerate(h)}
    amap = {}
    if obj:
        tmp = "/tmp/_summarize_cubin"
        os.makedirs(tmp, exist_ok=True)
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
        cubins = [f for f in os.listdir(tmp) if f.endswith(".cubin")]
        kern = rows[0][1].split("(")[0].split("::")[-1]
        for cb in cubins:
            txt = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cb)], capture_output=True, text=True).stdout
            cur, func = None, None
            for ln in txt.split("\n"):
                if ln.startswith(".text."):
                    func = ln
                    continue
                m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
                if m:
                    cur = "%s:%s" % (os.path.basename(m.group(1)), m.group(2))
                    continue
                m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", ln)
                if m and func and kern in func:
                    amap[int(m.group(1), 16)] = cur
            os.remove(os.path.join(tmp, cb))
    base = int(rows[2][0], 16)
    stcols = [k for k in h if k.startswith("stall_") and "Not Issued" not in k]
    byline = collections.Counter()
    linestall = collections.defaultdict(collections.Counter)
    bystall = collections.Counter()
    total = 0
    for r in rows[2:]:
        try:
            a = int(r[0], 16) - base
        except ValueError:
            continue
        n = int(r[ix["# Samples"]] or 0)
        total += n
        key = amap.get(a, "0x%05x" % (a // 0x400 * 0x400))
        byline[key] += n
        for k in stcols:
            v = int(r[ix[k]] or 0)
            bystall[k] += v
            linestall[key][k] += v
    with open(name + ".hotlines.txt", "w") as fh:
        fh.write("%s\nwarp-stall samples: %d\n\nby reason:\n" % (rows[0][1], total))
        for k, v in bystall.most_common(12):
            fh.write("  %-24s %9d  %5.1f %%\n" % (k, v, 100.0 * v / max(total, 1)))
        fh.write("\nby source line (top 40; stall_barrier = a warp waiting for the other roles of its CTA):\n")
        for k, v in byline.most_common(40):
            top = ", ".join("%s %d" % (a.replace("stall_", ""), b) for a, b in linestall[k].most_common(3))
            fh.write("  %-28s %9d  %5.1f %%   %s\n" % (k, v, 100.0 * v / max(total, 1), top))


if __name__ == "__main__":
    main()

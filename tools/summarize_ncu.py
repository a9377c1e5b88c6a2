#!/usr/bin/env python
"""Turn ncu exports into the small text summaries committed under profiles/.

  python tools/summarize_ncu.py launches gpurun_out/launches.csv          > profiles/launches_rNN.txt
  python tools/summarize_ncu.py report   gpurun_out/prof.ncu-rep [kernel] > profiles/ncu_rNN_<kernel>.txt
"""
import collections
import csv
import io
import subprocess
import sys

RAW = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
       "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
       "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
       "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
       "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
       "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
       "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
       "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum", "sm__icc_request_hit_rate.pct",
       "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
       "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        nm = r[ki].split("(")[0]
        v = float(r[vi].replace(",", ""))
        a = agg.setdefault(nm, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (serialised, cold-cache: compare SHARES)")
    print("%-44s %6s %12s %8s %10s" % ("kernel", "n", "total ms", "share", "avg ms"))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-44s %6d %12.3f %8.3f %10.4f" % (k, v[0], v[1] / 1e6, v[1] / tot, v[1] / v[0] / 1e6))


def report(path, kernel=None):
    cmd = ["ncu", "-i", path, "--page", "raw", "--csv"]
    if kernel:
        cmd += ["--kernel-name", "regex:" + kernel]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print("# ncu --set full --clock-control none: %s" % path)
    names = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("\n== %s" % r[names][:100])
        for m in RAW:
            if m in hdr:
                i = hdr.index(m)
                print("  %-78s %16s %s" % (m, r[i], units[i]))
        for i, h in enumerate(hdr):
            if "issue_stalled" in h and h.endswith("per_warp_active.pct") and "not_issued" not in h:
                try:
                    if float(r[i]) >= 3.0:
                        print("  %-78s %16s %%" % (h.replace("smsp__warp_", ""), r[i]))
                except ValueError:
                    pass
    # source hot spots
    cmd = ["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"]
    if kernel:
        cmd += ["--kernel-name", "regex:" + kernel]
    out = subprocess.run(cmd, capture_output=True, text=True).stdout
    cur, hdr = None, None
    agg, src = collections.Counter(), {}
    for r in csv.reader(io.StringIO(out)):
        if len(r) == 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) > 2 and r[0] == "Line No":
            hdr = r
            si = r.index("# Samples")
        elif hdr and len(r) == len(hdr) and r[0] != "":
            try:
                agg[(cur, int(r[0]))] += int(r[si])
                src[(cur, int(r[0]))] = r[1]
            except ValueError:
                pass
    tot = sum(agg.values()) or 1
    print("\n== warp-state samples by CUDA source line (top 15 of %d samples, all profiled launches of the selection)" % tot)
    for k, v in agg.most_common(15):
        print("  %5.1f%%  %s:%d  %s" % (100.0 * v / tot, k[0], k[1], src[k].strip()[:90]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        report(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)

"""Summarise ncu artefacts brought back in gpurun_out/ into small text files under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv profiles/r01_x_launches.txt
  python tools/ncu_summary.py full gpurun_out/prof.ncu-rep profiles/r01_x_full.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__average_warp_latency_per_inst_issued.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__thread_inst_executed_per_inst_executed.ratio"] + [
        "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % r for r in (
            "wait", "long_scoreboard", "short_scoreboard", "barrier", "not_selected", "math_pipe_throttle",
            "branch_resolving", "dispatch_stall", "no_instruction", "mio_throttle", "lg_throttle")]


def short(n):
    n = re.sub(r"\(.*", "", n)
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"chb::|<?unnamed>::", "", n)
    return n


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    h = rows[hi]
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    data = [(short(r[ki]), float(r[vi].replace(",", ""))) for r in rows[hi + 1:] if len(r) > vi]
    idx = [i for i, d in enumerate(data) if d[0].startswith("fused_particles_k")]
    if len(idx) < 2:
        idx = [i for i, d in enumerate(data) if d[0].startswith("push_coords_k")]
    out = ["ncu --metrics gpu__time_duration.sum --clock-control none: %d launches in %s" % (len(data), src)]
    if len(idx) >= 2:
        seg = data[idx[-2]:idx[-1]]
        out.append("one full PIC step (between the last two %s launches): %d launches" % (data[idx[-1]][0].split("<")[0], len(seg)))
    else:
        seg = data
    agg = collections.OrderedDict()
    for n, t in seg:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += t
    tot = sum(v[1] for v in agg.values())
    out.append("sum of kernel durations: %.3f ms (cold-cache, serialised: compare shares)" % (tot / 1e6))
    out.append("%-58s %6s %10s %7s" % ("kernel", "count", "ms", "share"))
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("%-58s %6d %10.3f %6.1f%%" % (n[:58], c, t / 1e6, 100 * t / tot))
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


def full(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    out = ["ncu --set full --clock-control none summary of %s" % src]
    for r in rows[2:]:
        out.append("== %s  (ID %s)" % (short(r[h.index("Kernel Name")]), r[h.index("ID")]))
        for k in KEYS:
            if k in h:
                i = h.index(k)
                out.append("   %-78s %s %s" % (k, r[i], units[i]))
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


def traffic(src, dst):
    """per-kernel DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum, mean over the captured
    launches) -> JSON consumed by bench.py for roofline.traffic"""
    import json

    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    h, units = rows[0], rows[1]
    ir, iw, it = h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum"), h.index("gpu__time_duration.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    tscale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
    agg = collections.OrderedDict()
    for r in rows[2:]:
        name = short(r[h.index("Kernel Name")])
        by = float(r[ir].replace(",", "")) * scale[units[ir]] + float(r[iw].replace(",", "")) * scale[units[iw]]
        a = agg.setdefault(name, [0, 0.0, 0.0])
        a[0] += 1
        a[1] += by
        a[2] += float(r[it].replace(",", "")) * tscale[units[it]]
    out = {"source": src, "note": "ncu --set full --clock-control none; mean per launch over the captured launches",
           "kernels": {k: {"launches": v[0], "dram_bytes_per_launch": v[1] / v[0], "ms_per_launch_under_ncu": v[2] / v[0]}
                       for k, v in agg.items()}}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic}[sys.argv[1]](sys.argv[2], sys.argv[3])

#!/usr/bin/env python
"""Turn the ncu captures brought back in gpurun_out/ into the text summaries kept under profiles/.

usage: make_profiles.py <tag> <launch-list csv> <k_hits .ncu-rep> <bench json> [<reference-arm json>]
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
KERNEL = "_ZN2nb6k_hitsENS_9RunParamsEPjNS_8HitQueueE"


def nm(s):
    return s.split("(")[0].replace("nb::", "").replace("<unnamed>::", "").replace("void ", "")[:44]


def main():
    tag, launches, rep, bench = sys.argv[1:5]
    refarm = sys.argv[5] if len(sys.argv) > 5 else None
    out = os.path.join(ROOT, "profiles")
    rows = [r for r in csv.reader(open(launches)) if len(r) > 14 and r[0].isdigit()]
    agg = collections.OrderedDict()
    for r in rows:
        a = agg.setdefault(nm(r[4]), [0, 0.0])
        a[0] += 1
        a[1] += float(r[14]) / 1e6
    tot = sum(v[1] for v in agg.values())
    b = json.load(open(bench))
    with open(os.path.join(out, "%s_bench_launch_list.txt" % tag), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline   (B200)\n")
        f.write("all %d launches of the bench command (cold-cache, serialised: compare SHARES, not absolutes)\n\n" % len(rows))
        f.write("%-46s %8s %12s %8s\n" % ("kernel", "launches", "total ms", "share"))
        for k, v in agg.items():
            f.write("%-46s %8d %12.3f %7.1f%%\n" % (k, v[0], v[1], 100 * v[1] / tot))
        ck = [k for k in ("k_pre", "k_hits", "k_hits_finish", "k_post") if k in agg]
        chain = sum(agg[k][1] for k in ck)
        f.write("\nchain kernels only (ncu): " + ", ".join("%s %.1f%%" % (k, 100 * agg[k][1] / chain) for k in ck) + "\n")
        s = b["roofline"]["stage_ms_per_step"]
        t = sum(s.values())
        f.write("bench.py CUDA-event stage shares (same build, no profiler): " + ", ".join("%s %.1f%%" % (k, 100 * v / t) for k, v in s.items()) + "\n")
        f.write("\nper-launch list, first 45 rows (id, kernel, block, grid, ms)\n")
        for r in rows[:45]:
            f.write("%s,%s,%s,%s,%.3f\n" % (r[0], nm(r[4]), r[7], r[8], float(r[14]) / 1e6))
    json.dump(b, open(os.path.join(out, "%s_bench_n1.json" % tag), "w"), indent=1)
    if refarm:
        json.dump(json.load(open(refarm)), open(os.path.join(out, "%s_bench_reference_arm.json" % tag), "w"), indent=1)
    kernels = [("k_hits", KERNEL, rep)]
    for short, mangled in (("k_pre", "_ZN2nb5k_preENS_9RunParamsEPi"), ("k_post", "_ZN2nb6k_postENS_9RunParamsEPi"),
                           ("k_hits_finish", "_ZN2nb13k_hits_finishENS_9RunParamsENS_8HitQueueE"),
                           ("k_lar", "_ZN2nb5k_larILi1EEEvNS_9LArParamsE")):
        cand = rep.replace("k_hits.", short + ".")
        if cand != rep and os.path.exists(cand):
            kernels.append((short, mangled, cand))
    sassd = os.path.join(ROOT, "gpurun_out", "sass")
    os.makedirs(sassd, exist_ok=True)
    for fn in os.listdir(sassd):
        if fn.endswith(".cubin"):
            os.remove(os.path.join(sassd, fn))
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "nest_b200", "csrc", "libnest_b200.so")], cwd=sassd,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    cub = [fn for fn in os.listdir(sassd) if fn.endswith(".cubin")][0]
    open(os.path.join(sassd, "all.sass"), "w").write(
        subprocess.run(["nvdisasm", "-g", cub], cwd=sassd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout)
    want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
            "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "dram__bytes_read.sum",
            "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
            "launch__shared_mem_per_block_static", "launch__shared_mem_per_block_dynamic", "local_load_store__"]
    for short, mangled, krep in kernels:
        raw = subprocess.run(["ncu", "-i", krep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
        rr = list(csv.reader(raw.splitlines()))
        hdr, units, vals = rr[0], rr[1], rr[2]
        with open(os.path.join(out, "%s_%s_ncu_full.txt" % (tag, short)), "w") as f:
            f.write("ncu --set full --clock-control none --import-source on -k regex:^%s$ -s 1 -c 1 python tools/profile_run.py --events 4e6 --launches 2   (B200)\n" % short)
            f.write("one launch of %s: 4e6 events of the bench workload (NR 0-100 keV flat, LUX_Run03), ~3.2e8 S1 PMT hits%s\n\n" % (short, "" if short != "k_lar" else " -- k_lar: tools/profile_lar.py, 2e7 ER events with fluctuations"))
            for h, u, v in zip(hdr, units, vals):
                if h in want or ("issue_stalled" in h and "per_issue_active" in h) or h.startswith("smsp__inst_executed_op_local") or h.startswith("l1tex__t_bytes_pipe_lsu_mem_local"):
                    f.write("%-92s %-16s %s\n" % (h, u, v))
        src = os.path.join(ROOT, "gpurun_out", "_src_%s_%s.csv" % (tag, short))
        open(src, "w").write(subprocess.run(["ncu", "-i", krep, "--page", "source", "--csv"], stdout=subprocess.PIPE, text=True).stdout)
        if short == "k_hits":  # everything inlined: attribute by inlined function and source line
            r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_line.py"), src,
                                os.path.join(sassd, "all.sass"), mangled, "24"], stdout=subprocess.PIPE, text=True).stdout
            open(os.path.join(out, "%s_%s_by_function.txt" % (tag, short)), "w").write(r)
        # out-of-line subroutines (binomial, philox_block, box_muller, libm helpers) kept apart
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_by_label.py"), krep, mangled, "30"],
                           stdout=subprocess.PIPE, text=True).stdout
        open(os.path.join(out, "%s_%s_by_subroutine.txt" % (tag, short)), "w").write(r)
    print(open(os.path.join(out, "%s_bench_launch_list.txt" % tag)).read()[:900])


if __name__ == "__main__":
    main()

"""Turn ncu captures under gpurun_out/ into small committed summaries under profiles/.
  python tools/summarize_profiles.py launches gpurun_out/launches_r1.csv profiles/r1_launches.md
  python tools/summarize_profiles.py ncu gpurun_out/prof_x.ncu-rep profiles/r1_ncu_x.md
"""
import collections, csv, re, subprocess, sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg, tot, n = collections.OrderedDict(), 0.0, 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:80]
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(row["Metric Unit"], 1)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v; n += 1
    with open(dst, "w") as f:
        f.write(f"# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised — compare SHARES)\n\n")
        f.write(f"source: `{src}` — {n} launches, {tot / 1e6:.2f} ms total\n\n| ms | share | launches | kernel |\n|---:|---:|---:|---|\n")
        for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
            f.write(f"| {t / 1e6:.3f} | {100 * t / tot:.1f}% | {c} | `{k}` |\n")
    print(open(dst).read())


def ncu(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of `{src}`\n\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
            f.write(f"## {d.get('Kernel Name', '?')[:120]}\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in d and d[k] != "":
                    f.write(f"| {k} | {d[k]} | {u.get(k, '')} |\n")
            f.write("\n")
    print(open(dst).read()[:3000])


if __name__ == "__main__":
    {"launches": launches, "ncu": ncu}[sys.argv[1]](sys.argv[2], sys.argv[3])

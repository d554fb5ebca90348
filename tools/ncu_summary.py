"""Summarise ncu output for profiles/: (1) a launch list (`ncu --metrics gpu__time_duration.sum --csv`) per kernel,
(2) selected metrics of a `--set full` report (.ncu-rep read with `ncu -i ... --page raw --csv`).

    python tools/ncu_summary.py launches gpurun_out/x_launches.csv
    python tools/ncu_summary.py report gpurun_out/x.ncu-rep
"""
import collections, csv, io, subprocess, sys

METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum",
           "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
           "launch__block_size", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
           "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem"]


def launches(path):
    rows = [r for r in csv.reader(l for l in open(path, errors="replace") if l.startswith('"'))]
    hdr = rows[0]
    ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= iv or r[im] != "gpu__time_duration.sum":
            continue
        a = agg.setdefault(r[ik][:90], [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        unit = 1e3 if tot > 1e7 else 1.0          # ns or us
        print(f"| `{k}` | {c} | {t / unit:.1f} | {t / unit / c:.1f} | {100 * t / tot:.1f} % |")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[0]
    print("| metric | " + " | ".join(f"launch {i + 1}" for i in range(len(rows) - 2)) + " |")
    print("|---|" + "---|" * (len(rows) - 2))
    print("| kernel | " + " | ".join(r[hdr.index("Kernel Name")][:60] for r in rows[2:]) + " |")
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            print(f"| {m} ({rows[1][i]}) | " + " | ".join(r[i] for r in rows[2:]) + " |")


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])

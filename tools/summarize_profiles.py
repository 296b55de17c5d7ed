"""Turn the raw ncu outputs a GPU run leaves in gpurun_out/ into the markdown summaries committed under profiles/.

    python tools/summarize_profiles.py launches gpurun_out/launches_cfg5.csv  profiles/r01b_launches_cfg5_summary.md  "title"
    python tools/summarize_profiles.py full     gpurun_out/prof_cfg5_raw.csv  profiles/r01b_ncu_full_cfg5_summary.md "title"
"""
import collections
import csv
import sys

METRICS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
           "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
           "l1tex__t_sector_hit_rate.pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "smsp__inst_executed.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
           "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_ltcfabric.sum",
           "smsp__pcsamp_warps_issue_stalled_long_scoreboard", "smsp__pcsamp_warps_issue_stalled_lg_throttle",
           "smsp__pcsamp_warps_issue_stalled_wait", "smsp__pcsamp_warps_issue_stalled_short_scoreboard"]


def rows_of(path):
    with open(path, newline="") as f:
        return [r for r in csv.reader(f) if len(r) > 10]


def launches(src, dst, title):
    rows = rows_of(src)
    h = rows[0]
    ik, im, iv = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
    tot = collections.OrderedDict()
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        t = float(r[iv].replace(",", "")) / 1e3    # ns -> us
        n, s = tot.get(r[ik], (0, 0.0))
        tot[r[ik]] = (n + 1, s + t)
    total = sum(s for _, s in tot.values())
    with open(dst, "w") as f:
        f.write(f"# {title}\n\n`ncu --metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES, "
                f"not absolute times).  {sum(n for n, _ in tot.values())} launches, {total / 1e3:.1f} ms in total.\n\n")
        f.write("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|\n")
        for k, (n, s) in sorted(tot.items(), key=lambda kv: -kv[1][1])[:30]:
            f.write(f"| `{k[:70]}` | {n} | {s:.1f} | {s / n:.1f} | {100 * s / total:.1f}% |\n")


def full(src, dst, title):
    rows = rows_of(src)
    h, units = rows[0], rows[1]
    ik = h.index("Kernel Name")
    body = rows[2:]
    cols = []
    for m in METRICS:
        cand = [i for i, name in enumerate(h) if name == m or name.endswith("." + m)]
        cand = [i for i in cand if body and body[0][i] != ""] or cand     # sections repeat names with empty cells
        if cand:
            cols.append((m, cand[0]))
    names = [f"{r[ik].split('(')[0].split('::')[-1].split('<')[0]}#{n + 1}" for n, r in enumerate(body)]
    with open(dst, "w") as f:
        f.write(f"# {title}\n\n`ncu --set full --clock-control none --import-source on`; values per launch from `ncu -i ... --page raw --csv`.\n\n")
        f.write("| metric | " + " | ".join(names) + " |\n|---|" + "---|" * len(names) + "\n")
        for m, i in cols:
            f.write(f"| {m} [{units[i]}] | " + " | ".join(r[i] for r in body) + " |\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3], sys.argv[4])

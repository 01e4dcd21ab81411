"""Developer tool: turn ncu outputs into the text summaries committed under profiles/.
  ncu_summary.py raw  <report.ncu-rep>      key metrics of the first kernel of an `ncu --set full` report
  ncu_summary.py list <launches.csv>        per-kernel launch counts / total time / share of an `ncu --metrics gpu__time_duration.sum --csv` log"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size",
    "launch__block_size", "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "sm__cycles_elapsed.avg.per_second",
]


def raw(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        print("# kernel: %s" % d.get("Kernel Name", ("", "?"))[1])
        for k in KEYS:
            if k in d:
                print("%-90s %-16s %s" % (k, d[k][0], d[k][1]))


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.reader(lines))
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.OrderedDict(), collections.Counter()
    for r in rows[1:]:
        if len(r) <= iv:
            continue
        name = r[ik].split("(")[0].replace("snarkv::", "").replace("void ", "")
        v = float(r[iv].replace(",", ""))
        u = r[iu]
        ms = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v if u in ("ms", "msecond") else v * 1e3
        tot[name] = tot.get(name, 0.0) + ms
        cnt[name] += 1
    pipeline = sum(v for k, v in tot.items() if k.startswith("k_") and not k.startswith("k_synth"))
    print("%-36s %8s %12s %8s" % ("kernel", "launches", "total_ms", "share"))
    for k, v in tot.items():
        share = "%.4f" % (v / pipeline) if k.startswith("k_") and not k.startswith("k_synth") else "-"
        print("%-36s %8d %12.3f %8s" % (k[:36], cnt[k], v, share))


if __name__ == "__main__":
    {"raw": raw, "list": launches}[sys.argv[1]](sys.argv[2])

"""Summarises ncu reports into the tables committed under profiles/ (run here, no GPU):
  python scripts/ncu_summary.py launches gpurun_out/launches_r02b.csv      -> per-kernel launch count / time / share
  python scripts/ncu_summary.py full gpurun_out/prof_r02b_cond.ncu-rep ...  -> key metrics per captured launch (markdown)
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem"),
    ("sm__cycles_elapsed.max", "SM cycles"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 read sectors (from SMs)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe (hmma) active %"),
    ("sm__inst_executed_pipe_tmem.sum", "TMEM instructions"),
    ("sm__inst_executed.sum", "warp instructions"),
    ("smsp__inst_executed.avg.per_cycle_active", "IPC per scheduler"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe active %"),
    ("sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "FMA-heavy pipe active %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "stall long scoreboard"),
    ("smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio", "stall short scoreboard"),
    ("smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio", "stall math pipe throttle"),
    ("smsp__average_warp_latency_issue_stalled_barrier.ratio", "stall barrier"),
    ("smsp__average_warp_latency_issue_stalled_wait.ratio", "stall wait"),
    ("smsp__average_warp_latency_issue_stalled_not_selected.ratio", "stall not selected"),
]


def launches(path):
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = OrderedDict()
    for r in rows[1:]:
        name = re.sub(r"\(.*", "", r[kn]).strip()
        name = re.sub(r"^void ", "", name)
        us = float(r[mv].replace(",", "")) / 1e3 if "ns" in r[hdr.index("Metric Unit")] else float(r[mv].replace(",", ""))
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print("kernel,launches,total_us,share_pct,avg_us")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%s,%d,%.1f,%.2f,%.2f" % (name.replace(",", ";"), n, t, 100 * t / tot, t / n))


def full(paths):
    for p in paths:
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        hdr, units = rows[0], rows[1]
        print("### `%s`\n" % p.split("/")[-1])
        names = [re.sub(r"^void ", "", re.sub(r"\(.*", "", r[hdr.index("Kernel Name")])) for r in rows[2:]]
        print("| metric | " + " | ".join("%s #%d" % (n.replace("fs::", ""), i) for i, n in enumerate(names)) + " |")
        print("|---|" + "---|" * len(names))
        for key, label in KEYS:
            if key not in hdr:
                continue
            i = hdr.index(key)
            vals = []
            for r in rows[2:]:
                try:
                    v = float(r[i].replace(",", ""))
                    vals.append(("%.4g" % v) + (" " + units[i] if units[i] else ""))
                except ValueError:
                    vals.append(r[i])
            print("| %s (`%s`) | " % (label, key) + " | ".join(vals) + " |")
        print()


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2:])

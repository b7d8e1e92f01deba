"""Summarises an ncu launch list (csv) and a full-set report (.ncu-rep) into markdown for profiles/.

    python tools/ncu_summary.py launches.csv report.ncu-rep > profiles/rN_summary.md
"""
import collections
import csv
import subprocess
import sys


def launch_list(path):
  rows = [r for r in csv.reader(open(path)) if len(r) > 5]
  hdr = next(r for r in rows if "Kernel Name" in r)
  data = rows[rows.index(hdr) + 1:]
  ki, vi, ui, mi = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Unit", "Metric Name"))
  agg = collections.OrderedDict()
  for r in data:
    if r[mi] != "gpu__time_duration.sum":
      continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
    name = r[ki].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    n, t = agg.get(name, (0, 0.0))
    agg[name] = (n + 1, t + v)
  total = sum(t for _, t in agg.values())
  print("| kernel | launches | total us | share | avg us |")
  print("|---|---|---|---|---|")
  for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.1f | %.1f%% | %.1f |" % (k, n, t, 100 * t / total, t / n))


WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "l1tex__t_sector_hit_rate.pct",
    "l1tex__average_t_sectors_per_request_pipe_lsu_mem_global_op_ld.ratio",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__issue_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__occupancy_limit_registers",
    "sm__cycles_elapsed.avg.per_second",
]


def full_report(path):
  out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
  rows = list(csv.reader(out.splitlines()))
  hdr, units, data = rows[0], rows[1], rows[2:]
  names = [r[hdr.index("Kernel Name")].split("(")[0].replace("void <unnamed>::", "") for r in data]
  print("| metric | unit | " + " | ".join("launch %d" % i for i in range(len(data))) + " |")
  print("|---|---|" + "---|" * len(data))
  for w in WANT:
    if w in hdr:
      i = hdr.index(w)
      print("| %s | %s | %s |" % (w, units[i], " | ".join(r[i] for r in data)))
  print("\nkernels: " + ", ".join(names))


if __name__ == "__main__":
  print("### launch list (`ncu --metrics gpu__time_duration.sum --clock-control none`)\n")
  launch_list(sys.argv[1])
  if len(sys.argv) > 2:
    print("\n### `ncu --set full --clock-control none` of `k_half_sweep<8>` "
          "(launches alternate node half / edge half)\n")
    full_report(sys.argv[2])

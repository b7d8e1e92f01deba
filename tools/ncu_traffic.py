"""Writes profiles/<key>_traffic.json from an `ncu --set full` report: DRAM bytes per launch of
the kernels whose name matches a pattern, stamped with the hash of the CUDA sources in the tree
(bench.py quotes the figure only while that hash still matches).

    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep c2 [--key half_sweep] [--kernel k_sweep] \
        [--source "profiles/r2_....md"]
"""
import argparse
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def to_bytes(value, unit):
  scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
  return float(value.replace(",", "")) * scale[unit]


def main():
  ap = argparse.ArgumentParser()
  ap.add_argument("report")
  ap.add_argument("workload")
  ap.add_argument("--key", default="half_sweep")
  ap.add_argument("--kernel", default="k_sweep")
  ap.add_argument("--source", default=None)
  args = ap.parse_args()
  from bench import kernel_source_sha
  out = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True,
                       text=True, check=True).stdout
  rows = list(csv.reader(out.splitlines()))
  hdr, units, data = rows[0], rows[1], rows[2:]
  ki = hdr.index("Kernel Name")
  ri, wi = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
  ti = hdr.index("gpu__time_duration.sum")
  launches = []
  for r in data:
    if args.kernel not in r[ki]:
      continue
    launches.append({"kernel": r[ki].split("(")[0].replace("void <unnamed>::", ""),
                     "dram_bytes_read": to_bytes(r[ri], units[ri]),
                     "dram_bytes_write": to_bytes(r[wi], units[wi]),
                     "duration_us_under_ncu": float(r[ti].replace(",", "")) *
                     {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(units[ti], 1.0)})
  if not launches:
    raise SystemExit("no launch of %r in %s" % (args.kernel, args.report))
  per_launch = sum(l["dram_bytes_read"] + l["dram_bytes_write"] for l in launches) / len(launches)
  doc = {"workload": args.workload, "kernel_pattern": args.kernel,
         "kernel_source_sha": kernel_source_sha(),
         "source": args.source or ("ncu --set full --clock-control none, %s" % args.report),
         "launches": launches, "bytes_per_launch": per_launch}
  path = os.path.join(ROOT, "profiles", args.key + "_traffic.json")
  json.dump(doc, open(path, "w"), indent=1)
  print("wrote %s: %.1f MB per launch over %d launches" % (path, per_launch / 1e6, len(launches)))


if __name__ == "__main__":
  main()

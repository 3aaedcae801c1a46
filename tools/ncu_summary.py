#!/usr/bin/env python
"""Turn an ncu capture into the two things the repo tracks: a text table under profiles/ and entries of
profiles/ncu_traffic.json (DRAM bytes per launch, which bench.py reports as `roofline.traffic`).

  # on the GPU box, next to the capture (reports are large, the pull-back limit is 64 MiB):
  ncu -i X.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\\
      sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,\\
      sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__inst_executed.sum,\\
      l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,lts__t_sector_hit_rate.pct > X_summary.csv
  # here:
  python tools/ncu_summary.py X_summary.csv --elements 100136441 --workload "100 MB text at -9, 112 chunks" \\
      --table profiles/r02_ncu_other_kernels.txt [--title "..."]

Only the first (largest) launch of every kernel instantiation becomes a traffic entry; k_text_pass2<...>
instantiations share the kernel name `k_text_pass2` (bench.py looks the newest matching entry up by kernel
name and element count)."""
import argparse
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
TIME = {"ns": 1e-3, "us": 1, "ms": 1e3, "s": 1e6}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("summary_csv")
    ap.add_argument("--elements", type=int, required=True, help="elements one launch processes (rotations, bytes, blocks)")
    ap.add_argument("--workload", required=True)
    ap.add_argument("--table", default=None, help="text file under profiles/ to append the table to")
    ap.add_argument("--title", default=None)
    ap.add_argument("--json", default=os.path.join(ROOT, "profiles", "ncu_traffic.json"))
    a = ap.parse_args()
    rows = list(csv.reader(open(a.summary_csv)))
    hdr, units = rows[0], dict(zip(rows[0], rows[1]))
    out = ["## " + (a.title or os.path.basename(a.summary_csv).replace("_summary.csv", "")),
           "%-44s %10s %10s %10s %7s %7s %7s %5s %9s %7s" % ("kernel (grid, block)", "time us", "dram rd MB", "dram wr MB", "sm %",
                                                             "issue %", "warps %", "regs", "warp-inst M", "L2 hit %")]
    traffic = json.load(open(a.json)) if os.path.exists(a.json) else []
    order = max([t.get("order", 0) for t in traffic] + [0])
    seen = set()
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(")[0].replace("void ", "")
        t = float(d["gpu__time_duration.sum"]) * TIME.get(units["gpu__time_duration.sum"], 1)
        rd = float(d["dram__bytes_read.sum"]) * SCALE.get(units["dram__bytes_read.sum"], 1)
        wr = float(d["dram__bytes_write.sum"]) * SCALE.get(units["dram__bytes_write.sum"], 1)
        out.append("%-44s %10.1f %10.1f %10.1f %7.1f %7.1f %7.1f %5s %9.1f %7.1f" % (
            (name + " " + d["Grid Size"] + " " + d["Block Size"])[:44], t, rd / 1e6, wr / 1e6,
            float(d["sm__throughput.avg.pct_of_peak_sustained_elapsed"]), float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
            float(d["sm__warps_active.avg.pct_of_peak_sustained_active"]), d["launch__registers_per_thread"],
            float(d["smsp__inst_executed.sum"]) / 1e6, float(d["lts__t_sector_hit_rate.pct"])))
        if name not in seen:
            seen.add(name)
            order += 1
            traffic.append({"order": order, "kernel": name.split("<")[0], "instance": name, "elements": a.elements,
                            "dram_bytes": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr), "duration_us": round(t, 1),
                            "source": a.table or a.summary_csv, "workload": a.workload})
    text = "\n".join(out) + "\n"
    print(text)
    if a.table:
        with open(a.table, "a") as f:
            f.write(text)
    json.dump(traffic, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()

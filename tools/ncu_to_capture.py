#!/usr/bin/env python
"""Turn `ncu --set full` reports into the capture list bench.py reads for `roofline.traffic`
(profiles/r02_train_kernel.json) and a short metric table (markdown).

    python tools/ncu_to_capture.py gpurun_out/r02_train_p1.ncu-rep:1 gpurun_out/r02_train_p8.ncu-rep:8 > profiles/r02_train_kernel.md
(report:num_partition; the JSON is written next to the markdown as profiles/r02_train_kernel.json)"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "lts__t_sectors.avg", "lts__t_sectors.max", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def raw(report):
    text = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
    rows = list(csv.reader(io.StringIO(text)))
    header, units, values = rows[0], rows[1], rows[2]
    return dict(zip(header, values)), dict(zip(header, units))


def main():
    git = subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, stdout=subprocess.PIPE, text=True).stdout.strip()
    captures, tables = [], []
    for item in sys.argv[1:]:
        report, _, partition = item.partition(":")
        values, units = raw(report)
        kernel = values["Kernel Name"]
        name = kernel.split("(")[0]
        if "<" in name:
            base, args = name.split("<", 1)
            first = args.split(",")[0].strip(" >")
            name = "gv::device::%s<%s>" % (base.split("::")[-1].replace("void ", "").strip(), first)
        dram = sum(float(values[k]) * SCALE[units[k]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        edges = int(values["launch__grid_size"]) * int(values["launch__block_size"]) // 32
        captures.append({"kernel": name, "num_partition": int(partition or 1), "edges_per_launch": edges,
                         "dram_bytes_per_launch": dram, "git": git, "report": os.path.basename(report),
                         "source": "ncu --set full --clock-control none, one warm launch of `python bench.py --steps 2 "
                                   "--warmup 1 --no-e2e --no-cpu-baseline --partitions %s`" % (partition or 1)})
        tables.append((item, kernel, values, units))
    with open(os.path.join(ROOT, "profiles", "r02_train_kernel.json"), "w") as fout:
        json.dump({"captures": captures}, fout, indent=1)
    for item, kernel, values, units in tables:
        print("### %s\n\n`%s`\n\n| metric | value |\n|---|---|" % (item, kernel[:120]))
        for key in KEYS:
            if key in values:
                print("| %s | %s %s |" % (key, values[key], units[key]))
        print()


if __name__ == "__main__":
    main()

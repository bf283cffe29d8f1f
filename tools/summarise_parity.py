#!/usr/bin/env python
"""Markdown tables of the parity sweeps committed under profiles/ (one JSON object per run / per summary, written by
tools/parity_sweep.py on a B200):  python tools/summarise_parity.py profiles/r02_parity*.jsonl > profiles/r02_parity_tables.md"""
import json
import os
import sys


def main():
    for path in sys.argv[1:]:
        header, reference, rows = None, None, []
        for line in open(path):
            row = json.loads(line)
            if "workload" in row and "summary" not in row and "impl" not in row:
                header = row
            elif row.get("summary") == "reference":
                reference = row
            elif "summary" in row:
                rows.append(row)
        if not rows:
            continue
        print("### %s -- %s, %d epochs%s\n" % (os.path.basename(path), header["workload"], header["epochs"],
                                                ", num_partition %d" % header["partitions"] if header.get("partitions") else ""))
        if reference:
            print("reference (unmodified, %d runs): |vertex| %.2f (range %.2f = %.2e rel), |context| %.2f (range %.2f = "
                  "%.2e rel), AUC %.4f (range %.4f)\n" %
                  (reference["runs"], reference["vertex_norm"], reference["vertex_norm_spread"],
                   reference["vertex_norm_spread"] / reference["vertex_norm"], reference["context_norm"],
                   reference["context_norm_spread"], reference["context_norm_spread"] / reference["context_norm"],
                   reference["auc"], reference["auc_spread"]))
        print("| setting | runs | |vertex| rel. diff | |context| rel. diff | AUC diff | AUC range | kernel edges/s | "
              "train loop edges/s |")
        print("|---|---|---|---|---|---|---|---|")
        for r in rows:
            print("| `%s` | %d | %s | %s | %s | %.4f | %.3g | %.3g |" %
                  (r["summary"], r["runs"],
                   "%+.2f %%" % (100 * r["vertex_norm_rel"]) if "vertex_norm_rel" in r else "-",
                   "%+.2f %%" % (100 * r["context_norm_rel"]) if "context_norm_rel" in r else "-",
                   "%+.4f" % r["auc_diff"] if "auc_diff" in r else "-", r["auc_spread"],
                   r.get("kernel_edges_per_s", 0), r.get("train_loop_edges_per_s", 0)))
        print()


if __name__ == "__main__":
    main()

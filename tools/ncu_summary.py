"""Summarise an `ncu --page raw --csv` export into the two files kept under profiles/.

    ncu -i prof.ncu-rep --page raw --csv > raw.csv
    python tools/ncu_summary.py raw.csv profiles/r1_v5

writes  <prefix>_ncu_full_summary.csv  (one row per launch, the columns the design discussion uses)
and     <prefix>_dram_traffic.json     (dram read+write bytes per ACTIVE launch: launches longer than 8 us;
                                        the no-op launches of a solve that ended early are excluded).
"""
import csv
import json
import sys

COLS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__cluster_size", "sm__cycles_elapsed.max", "smsp__inst_executed.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.sum", "smsp__inst_executed_pipe_fp64.sum",
]
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "nsecond": 1e-3, "usecond": 1.0, "us": 1.0,
         "ns": 1e-3, "msecond": 1e3, "ms": 1e3}


def main():
    raw, prefix = sys.argv[1], sys.argv[2]
    rows = list(csv.reader(open(raw)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    names, units = rows[hdr], rows[hdr + 1]
    col = {n: i for i, n in enumerate(names)}
    keep = [c for c in COLS if c in col]
    out, traffic = [], {}
    for r in rows[hdr + 2:]:
        if len(r) < len(names):
            continue
        kname = r[col["Kernel Name"]]
        vals = {}
        for c in keep:
            try:
                vals[c] = float(r[col[c]].replace(",", ""))
            except ValueError:
                vals[c] = float("nan")
        us = vals["gpu__time_duration.sum"] * SCALE.get(units[col["gpu__time_duration.sum"]], 1.0)
        dram = sum(vals[c] * SCALE.get(units[col[c]], 1.0) for c in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        out.append([r[col["ID"]], kname] + [r[col[c]] for c in keep])
        t = traffic.setdefault(kname, {"launches": 0, "active": 0, "_bytes": 0.0, "_us": 0.0})
        t["launches"] += 1
        if us > 8.0:
            t["active"] += 1
            t["_bytes"] += dram
            t["_us"] += us
    with open(prefix + "_ncu_full_summary.csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["ID", "Kernel Name"] + keep)
        w.writerow(["", ""] + [units[col[c]] for c in keep])
        w.writerows(out)
    res = {}
    for k, t in traffic.items():
        a = max(t["active"], 1)
        res[k] = {"launches": t["launches"], "active": t["active"], "dram_bytes_per_active_launch": t["_bytes"] / a,
                  "avg_us_active_under_ncu": t["_us"] / a}
    json.dump(res, open(prefix + "_dram_traffic.json", "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()

"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, share, mean).

    python profiles/summarize_launches.py gpurun_out/r01_launches_cfg3_v5.csv [first_launch_id last_launch_id]
"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hdr]
    ki, vi, mi, ii = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name"), h.index("ID")
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 60
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hdr + 1:]:
        if len(r) <= vi or r[mi] != "gpu__time_duration.sum" or not (lo <= int(r[ii]) <= hi):
            continue
        name = r[ki].split("(")[0].replace("void ", "").replace("gg::", "").replace("<unnamed>::", "")
        agg[name][0] += 1
        agg[name][1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':58s} {'n':>6s} {'total us':>10s} {'share':>6s} {'mean us':>8s}")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        print(f"{k[:58]:58s} {v[0]:6d} {v[1] / 1e3:10.1f} {v[1] / tot:6.3f} {v[1] / v[0] / 1e3:8.1f}")
    print(f"{'TOTAL':58s} {sum(v[0] for v in agg.values()):6d} {tot / 1e3:10.1f}")


if __name__ == "__main__":
    main()

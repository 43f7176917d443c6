"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel: count, total us, share.
Usage: python tools/launch_summary.py gpurun_out/launches.csv [--top N] [--per-launch]"""
import collections
import csv
import io
import re
import sys


def load(path):
    txt = open(path).read()
    start = txt.find('"ID"')
    return list(csv.DictReader(io.StringIO(txt[start:])))


def main():
    rows = load(sys.argv[1])
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        n = re.sub(r"\(.*", "", r["Kernel Name"])
        agg[n][0] += 1
        agg[n][1] += float(r["Metric Value"]) / 1e3
    tot = sum(v[1] for v in agg.values())
    print(f"launches {len(rows)}  total {tot:.1f} us (cold-cache, serialised: compare SHARES)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:36s} n={v[0]:5d} us={v[1]:11.1f} share={v[1] / tot:6.3f} avg_us={v[1] / v[0]:9.1f}")
    if "--per-launch" in sys.argv:
        top = sorted(rows, key=lambda r: -float(r["Metric Value"]))[:40]
        for r in top:
            print(r["ID"], re.sub(r"\(.*", "", r["Kernel Name"]), r["Grid Size"], r["Block Size"], float(r["Metric Value"]) / 1e3)


if __name__ == "__main__":
    main()

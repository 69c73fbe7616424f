#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time, count and
share per kernel name.  usage: launch_summary.py launches.csv [top_n] [skip_first_n_launches]"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    skip = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    hdr = None
    agg, cnt = collections.Counter(), collections.Counter()
    n = 0
    for r in csv.reader(open(path, errors="ignore")):
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d.get("Metric Name") != "gpu__time_duration.sum":
                continue
            n += 1
            if n <= skip:
                continue
            v = float(d["Metric Value"].replace(",", ""))
            u = d["Metric Unit"]
            v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
            k = d["Kernel Name"][:120]
            agg[k] += v
            cnt[k] += 1
    tot = sum(agg.values())
    print("# total kernel time %.1f us over %d launches" % (tot, sum(cnt.values())))
    print("#   time_us  count  share  kernel")
    for k, v in agg.most_common(top):
        print("%10.1f %5d %5.1f%% %s" % (v, cnt[k], 100 * v / tot, k))


if __name__ == "__main__":
    main()

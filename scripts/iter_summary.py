#!/usr/bin/env python
"""One-iteration kernel summary from an ncu launch list of `bench.py --no-graphs`
(iterations are delimited by the 4th adam_kernel launch).  usage: iter_summary.py csv [top]"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
    rows, hdr = [], None
    for r in csv.reader(open(path, errors="ignore")):
        if len(r) > 5 and r[0] == "ID":
            hdr = r
            continue
        if hdr and len(r) == len(hdr):
            d = dict(zip(hdr, r))
            if d["Metric Name"] == "gpu__time_duration.sum":
                v = float(d["Metric Value"].replace(",", ""))
                u = d["Metric Unit"]
                v = v / 1e3 if u == "ns" else (v * 1e3 if u == "ms" else v)
                rows.append((d["Kernel Name"], v))
    idx = [i for i, (k, v) in enumerate(rows) if "adam_kernel" in k]
    iters, start = [], 0
    for j in range(3, len(idx), 4):
        iters.append(rows[start:idx[j] + 1])
        start = idx[j] + 1
    last = iters[-2] if len(iters) >= 2 else iters[-1]
    agg, cnt = collections.Counter(), collections.Counter()
    for k, v in last:
        agg[k[:110]] += v
        cnt[k[:110]] += 1
    tot = sum(agg.values())
    print("# one iteration: total kernel time %.1f us over %d launches (%d iterations in the file)"
          % (tot, sum(cnt.values()), len(iters)))
    print("#   time_us  count  share  kernel")
    for k, v in agg.most_common(top):
        print("%10.1f %5d %5.1f%% %s" % (v, cnt[k], 100 * v / tot, k))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Coarse view of a graph_timeline.py TSV: per-stream spans and, per 250 us window, the kernels in flight.
usage: timeline_phases.py timeline.tsv [t0] [t1]"""
import collections
import sys

rows = [l.rstrip('\n').split('\t') for l in open(sys.argv[1])]
rows = [(float(a), float(b), c, d) for a, b, c, d in rows]
t0 = float(sys.argv[2]) if len(sys.argv) > 2 else 0
t1 = float(sys.argv[3]) if len(sys.argv) > 3 else 1e9
by = collections.defaultdict(list)
for s, d, st, n in rows:
    by[st].append((s, d, n))
for st, v in sorted(by.items(), key=lambda kv: kv[1][0][0]):
    print("stream %s: %d kernels, first %.0f last end %.0f, busy %.0f us" % (
        st, len(v), v[0][0], max(a + b for a, b, _ in v), sum(b for _, b, _ in v)))
short = lambda n: n.replace('void ', '').replace('(anonymous namespace)::', '').replace('at::native::', '')[:34]
w = 250
k = int(t0 // w)
while k * w < min(t1, max(r[0] + r[1] for r in rows)):
    a, b = k * w, (k + 1) * w
    sel = [r for r in rows if a <= r[0] < b]
    k += 1
    if not sel:
        continue
    c = collections.Counter()
    for s, d, st, n in sel:
        c[short(n)] += d
    print("%5d-%5d n=%3d busy=%5.0f nstreams=%d | %s" % (a, b, len(sel), sum(r[1] for r in sel), len(set(r[2] for r in sel)),
                                                     '; '.join('%s %.0f' % (x, y) for x, y in c.most_common(5))))

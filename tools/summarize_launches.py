#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections
import csv
import sys


def main(path, title=""):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
    H = rows[hdr]
    ik, iv, iu = H.index('Kernel Name'), H.index('Metric Value'), H.index('Metric Unit')
    ig = H.index('Grid Size') if 'Grid Size' in H else None
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hdr + 1:]:
        if len(r) <= iv:
            continue
        v = float(r[iv].replace(',', ''))
        v *= {'ns': 1.0, 'us': 1e3, 'ms': 1e6, 's': 1e9}.get(r[iu], 1.0)
        k = r[ik].split('(')[0]
        if ig is not None and 'k_gsf' in k:      # the march stages share one instantiation
            k += " grid=" + r[ig].replace(' ', '')
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    if title:
        print("# " + title)
    print("# cold-cache, serialised per-launch times: compare SHARES, not absolutes")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-90s launches=%6d  total=%10.3f ms  share=%5.1f%%"
              % (k[:90], v[0], v[1] / 1e6, 100 * v[1] / tot))


if __name__ == "__main__":
    main(sys.argv[1], " ".join(sys.argv[2:]))

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
    hdr = rows[hi]
    kn, mv = hdr.index('Kernel Name'), hdr.index('Metric Value')
    agg = collections.OrderedDict()
    n = 0
    for r in rows[hi + 1:]:
        if len(r) <= mv:
            continue
        name = r[kn].split('(')[0].replace('dmc::', '').replace('void ', '')
        t = float(r[mv].replace(',', '')) / 1000.0
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += t
        n += 1
    tot = sum(v[1] for v in agg.values())
    print('# %s: %d launches, %.1f us total (ncu per-launch times are cold-cache and serialised: compare shares)'
          % (path, n, tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-44s n=%3d  %9.1f us  %5.1f%%' % (k[:44], v[0], v[1], 100 * v[1] / tot))


if __name__ == '__main__':
    main(sys.argv[1])

"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel
for the window of one loglike+grad evaluation (from the second-to-last
gram_kernel launch of the value loop to the next one).
usage: summarize_launches.py launches.csv [first_gram_index] > summary.txt"""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    with open(path) as f:
        lines = [l for l in f if not l.startswith('==')]
    rows = []
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        v = float(row['Metric Value'].replace(',', ''))
        u = row['Metric Unit']
        us = {'ns': v/1e3, 'us': v, 'ms': v*1e3, 's': v*1e6}[u]
        name = row['Kernel Name']
        short = name.split('(')[0].replace('void ', '').replace('pgp::', '').replace('<unnamed>::', '')
        short = short.split('<')[0]
        rows.append((short, us))
    grams = [i for i, (n, _) in enumerate(rows) if n == 'gram_kernel']
    k = int(sys.argv[2]) if len(sys.argv) > 2 else len(grams) - 2
    lo, hi = grams[k], grams[k + 1]
    # the evaluation ends with trace_finish_kernel; cut the window there
    for j in range(lo, hi):
        if rows[j][0] == 'trace_finish_kernel':
            hi = j + 1
    # the scale_kernel launched just before the Gram build belongs to the evaluation
    if lo > 0 and rows[lo - 1][0] == 'scale_kernel':
        lo -= 1
    win = rows[lo:hi]
    agg = collections.OrderedDict()
    for n, us in win:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    print('# window = launches %d..%d of %s (one set_hyper + loglikelihood(True) evaluation)' % (lo, hi - 1, path))
    print('# per-launch times are cold-cache and serialised by ncu: compare SHARES with bench.py, not absolutes')
    print('%-28s %8s %12s %7s' % ('kernel', 'launches', 'total_us', 'share'))
    for n, a in sorted(agg.items(), key=lambda x: -x[1][1]):
        print('%-28s %8d %12.1f %6.1f%%' % (n, a[0], a[1], 100*a[1]/tot))
    print('%-28s %8d %12.1f' % ('TOTAL', len(win), tot))


if __name__ == '__main__':
    main()

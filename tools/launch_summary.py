"""Per-kernel summary of an ncu --csv launch list (metrics as rows): python tools/launch_summary.py file.csv [max_per_kernel]"""
import collections
import csv
import sys


def main(path, per=2):
    rows = list(csv.reader(open(path)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == 'ID')
    hdr, data = rows[hi], rows[hi + 1:]
    ix = {n: i for i, n in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in data:
        if len(r) < len(hdr):
            continue
        agg.setdefault((r[ix['ID']], r[ix['Kernel Name']][:44]), {})[r[ix['Metric Name']]] = r[ix['Metric Value']]
    seen = collections.Counter()
    for (id_, name), m in agg.items():
        seen[name] += 1
        if seen[name] > per:
            continue
        short = {k.split('.')[0].replace('gpu__', '').replace('smsp__', '').replace('sm__', ''): v for k, v in m.items()}
        print(id_, name, short)


if __name__ == '__main__':
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 2)

"""Summarise an ncu launch list (csv from --metrics gpu__time_duration.sum) for profiles/."""
import collections
import csv
import re
import sys


def load(path):
    lines = [l for l in open(path) if not l.startswith('==')]
    seq = []
    for row in csv.DictReader(lines):
        try:
            t = float(row['Metric Value'].replace(',', ''))
        except Exception:
            continue
        t *= {'ns': 1, 'us': 1e3, 'ms': 1e6}.get(row['Metric Unit'], 1)
        seq.append((row['Kernel Name'], t))
    return seq


def short(n):
    n = re.sub(r'\(.*', '', n)
    n = n.replace('void ', '').replace('rlsb::<unnamed>::', 'rlsb::')
    return n[:90]


if __name__ == '__main__':
    seq = load(sys.argv[1])
    lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    hi = int(sys.argv[3]) if len(sys.argv) > 3 else len(seq)
    sub = seq[lo:hi]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, t in sub:
        agg[short(n)][0] += 1
        agg[short(n)][1] += t
    tot = sum(v[1] for v in agg.values())
    mine = sum(v[1] for k, v in agg.items() if 'rlsb::' in k)
    print(f"launches {len(sub)}  total {tot/1e6:.3f} ms  (librlsb kernels {mine/1e6:.3f} ms = {100*mine/tot:.1f}%)")
    print("| kernel | launches | total ms | share |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        print(f"| `{k}` | {v[0]} | {v[1]/1e6:.3f} | {100*v[1]/tot:.1f}% |")

"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py: the launches of ONE resident step
(from a k_assign_columns launch to the next one) by kernel, with shares.  Usage: python tools/summarize_launches.py <csv> [step]"""
import csv, re, sys, collections

rows = []
for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')):
    if r[0] == "ID":
        continue
    name = re.sub(r"^(void )?h2b::", "", r[4])
    name = re.sub(r"\(.*$", "", name)
    rows.append((name, float(r[-1]) / 1e3))
marks = [i for i, (n, _) in enumerate(rows) if n.startswith("k_assign_columns")]
step = int(sys.argv[2]) if len(sys.argv) > 2 else 1
a, b = marks[step], marks[step + 1]
agg = collections.OrderedDict()
for n, us in rows[a:b]:
    c, t = agg.get(n, (0, 0.0))
    agg[n] = (c + 1, t + us)
tot = sum(t for _, t in agg.values())
print(f"launches {a}..{b - 1} of {len(rows)}: {b - a} launches, {tot:.1f} us serialised")
print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{n}` | {c} | {t:.1f} | {100 * t / tot:.1f}% |")
print(f"| **sum** | {b - a} | {tot:.1f} | 100% |")

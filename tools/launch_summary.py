"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel time of ONE step
(the last `--per-step` launches, or everything)."""
import collections, csv, sys

def main(path, per_step=None):
    rows = list(csv.reader(open(path, errors="replace")))
    h = next(r for r in rows if "Kernel Name" in r)
    ki, vi = h.index("Kernel Name"), h.index("Metric Value")
    data = []
    for r in rows[rows.index(h) + 1:]:
        if len(r) > vi:
            try:
                data.append((r[ki], float(r[vi].replace(",", ""))))
            except ValueError:
                pass
    if per_step:
        data = data[-per_step:]
    d = collections.defaultdict(lambda: [0, 0.0])
    for k, v in data:
        d[k[:100]][0] += 1
        d[k[:100]][1] += v
    tot = sum(v[1] for v in d.values())
    print(f"# {path}: {len(data)} launches, {tot / 1e3:.1f} us")
    for k, v in sorted(d.items(), key=lambda kv: -kv[1][1]):
        print(f"{v[1] / 1e3:10.1f} us {v[0]:5d}x {100 * v[1] / tot:5.1f}%  {k}")

if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else None)

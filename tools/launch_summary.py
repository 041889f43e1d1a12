"""Per-kernel summary of an ncu launch list (`ncu --metrics gpu__time_duration.sum --csv --log-file X`): launches, mean / total us, share.
Kernels launched through the one launch path carry the functor as first template argument (k_one<k_x, ...> / k_many<k_x, ...>).
Usage: python tools/launch_summary.py <launches.csv> [<out.csv>]"""
import collections, csv, re, sys
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
r = list(csv.reader(lines))
hdr = r[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for x in r[1:]:
    if len(x) <= vi:
        continue
    m = re.search(r"k_(?:one|many)<(?:b2::)?(k_[a-z0-9_]+(?:<[^>]*>)?)", x[ki]) or re.search(r"(k_[a-z0-9_]+)", x[ki])
    try:
        v = float(x[vi].replace(",", ""))
    except ValueError:
        continue
    agg.setdefault(m.group(1) if m else x[ki][:48], []).append(v)
tot = sum(sum(v) for v in agg.values())
out = [("kernel", "launches", "mean_us", "total_us", "share_pct")]
for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    out.append((k, len(v), round(sum(v) / len(v) / 1e3, 2), round(sum(v) / 1e3, 1), round(100 * sum(v) / tot, 1)))
w = csv.writer(open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout)
w.writerows(out)

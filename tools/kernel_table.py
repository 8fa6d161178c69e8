"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total ms, share.
usage: kernel_table.py launches.csv [divide_by]"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
div = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, cnt = collections.defaultdict(float), collections.Counter()
for r in rows[hi + 2:]:
    if len(r) <= mv:
        continue
    name = re.sub(r"\(.*", "", r[kn]).replace("void ", "").replace("bj::", "")
    name = re.sub(r"native::.*?(\w+)<.*", r"torch:\1", name)
    tot[name] += float(r[mv].replace(",", "")) / 1e6 / div
    cnt[name] += 1
T = sum(tot.values())
print("total kernel time %.1f ms, %d launches (cold-cache, serialised by ncu)" % (T, sum(cnt.values()) / div))
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    if v / T < 0.002:
        continue
    print("%9.2f ms %5.1f%%  x%-5d %s" % (v, 100 * v / T, cnt[k] / div, k[:100]))

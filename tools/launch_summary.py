"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
Only whole steps are counted: the window runs from the first to the last `dice_finish_kernel`
launch (one per training step), so the per-step figures do not depend on where the capture began.
usage: python tools/launch_summary.py gpurun_out/launches.csv > profiles/xxx.csv"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    marker = sys.argv[2] if len(sys.argv) > 2 else "dice_finish_kernel"
    rows = list(csv.reader(open(path, errors="replace")))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    idx = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[idx["Kernel Name"]]
        name = re.sub(r"^void ", "", name)
        name = re.sub(r"\(.*$", "", name)
        name = name.replace("nas3d::", "").replace("(anonymous namespace)::", "")
        unit = r[idx["Metric Unit"]]
        v = float(r[idx["Metric Value"]].replace(",", ""))
        us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
        launches.append((name, us))
    marks = [i for i, (n, _) in enumerate(launches) if marker in n]
    steps = 1.0
    if len(marks) >= 2:
        launches = launches[marks[0] + 1:marks[-1] + 1]
        steps = float(len(marks) - 1)
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for name, us in launches:
        agg[name][0] += 1
        agg[name][1] += us
        tot += us
    print("# whole steps in window: %d" % steps)
    print("kernel,launches_per_step,us_per_step,share")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('"%s",%.1f,%.1f,%.4f' % (k, n / steps, us / steps, us / tot))
    print('"TOTAL",%.1f,%.1f,1.0' % (sum(v[0] for v in agg.values()) / steps, tot / steps))


if __name__ == "__main__":
    main()

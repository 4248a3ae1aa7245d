"""Summarise ncu CSV output into markdown tables.

    python scripts/summarize_ncu.py launches <launches.csv>      # --metrics gpu__time_duration.sum launch list
    python scripts/summarize_ncu.py full <raw.csv>               # `ncu -i x.ncu-rep --page raw --csv` of a --set full capture
"""
import csv
import io
import re
import sys
from collections import OrderedDict


def rows(path):
    txt = open(path, errors="replace").read()
    start = txt.find('"ID"')
    return list(csv.DictReader(io.StringIO(txt[start:])))


def short(name):
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*$", "", name)[:90]


def launches(path):
    agg = OrderedDict()
    for r in rows(path):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        us = v / 1e3 if r["Metric Unit"] in ("ns", "nsecond") else (v if r["Metric Unit"] in ("us", "usecond") else v * 1e3)
        a = agg.setdefault(short(r["Kernel Name"]), [0, 0.0])
        a[0] += 1
        a[1] += us
    tot = sum(a[1] for a in agg.values())
    n = sum(a[0] for a in agg.values())
    print("| kernel | launches | total us | us/launch | share |\n|---|---:|---:|---:|---:|")
    for k, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f | %.1f%% |" % (k, c, t, t / c, 100 * t / tot))
    print("\ntotal %.1f us over %d launches" % (tot, n))


def full(path):
    want = OrderedDict([("gpu__time_duration.sum", "time us"), ("dram__bytes_read.sum", "DRAM read MB"), ("dram__bytes_write.sum", "DRAM write MB"),
                        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM % of peak"),
                        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"), ("launch__registers_per_thread", "regs"),
                        ("launch__grid_size", "grid")])
    rs = rows(path)
    agg = OrderedDict()
    for r in rs[1:] if rs and not rs[0].get("ID", "").isdigit() else rs:
        k = short(r["Kernel Name"]) + " " + r.get("Grid Size", r.get("launch__grid_size", "?"))
        a = agg.setdefault(k, dict(n=0, **{m: 0.0 for m in want}))
        a["n"] += 1
        for m in want:
            col = m if m in r else next((c for c in r if c.endswith(m)), None)   # section-prefixed columns
            try:
                a[m] += float(r[col].replace(",", ""))
            except (KeyError, ValueError, TypeError):
                pass
    units = rs[0] if rs and not rs[0].get("ID", "").isdigit() else {}
    print("| kernel (grid) | launches | " + " | ".join(list(want.values())[:-1]) + " |\n|---|---:|" + "---:|" * (len(want) - 1))
    for k, a in agg.items():
        n = a["n"]
        t = a["gpu__time_duration.sum"] / n
        t = t / 1e3 if units.get("gpu__time_duration.sum", "ns").startswith("n") else t
        scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}
        rd = a["dram__bytes_read.sum"] / n * scale.get(units.get("dram__bytes_read.sum", "byte"), 1e-6)
        wr = a["dram__bytes_write.sum"] / n * scale.get(units.get("dram__bytes_write.sum", "byte"), 1e-6)
        print("| `%s` | %d | %.1f | %.1f | %.1f | %.1f | %.1f | %d |" % (
            k, n, t, rd, wr, a["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"] / n,
            a["sm__warps_active.avg.pct_of_peak_sustained_active"] / n, a["launch__registers_per_thread"] / n))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])

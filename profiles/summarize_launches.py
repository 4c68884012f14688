#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, share, average)."""
import collections, csv, io, re, sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for r in csv.DictReader(io.StringIO("".join(lines))):
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        v = float(r["Metric Value"].replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}[r["Metric Unit"]]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if "rlx::" in k or "tc::tc_gemm" in k)
    print(f"# {path}: {sum(a[0] for a in agg.values())} launches, {tot / 1e6:.2f} ms total device time (cold-cache, serialised by ncu)")
    print(f"# rlx:: / rlx::tc:: kernels (this repo): {100 * ours / tot:.2f}% of device time; the rest are torch kernels of the synthetic env / metric plumbing")
    print(f"{'total ms':>10} {'share':>7} {'launches':>9} {'avg us':>10}  kernel")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{t / 1e6:10.2f} {100 * t / tot:6.2f}% {n:9d} {t / n / 1e3:10.2f}  {k[:140]}")


if __name__ == "__main__":
    main(sys.argv[1])

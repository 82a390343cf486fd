"""Aggregates an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel: launches, total time, share.
    python tools/launch_summary.py gpurun_out/launches.csv [title] > profiles/rNN_launches.md"""
import collections
import csv
import io
import re
import sys


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    name = name.replace("llb::", "").replace("(bool)", "").replace("(int)", "")
    return name


def main():
    path = sys.argv[1]
    title = sys.argv[2] if len(sys.argv) > 2 else path
    lines = open(path, errors="replace").read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    agg = collections.OrderedDict()
    order = []
    for row in csv.DictReader(io.StringIO("\n".join(lines[start:]))):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        ns = float(row["Metric Value"].replace(",", ""))
        if row.get("Metric Unit") == "us":
            ns *= 1e3
        k = short(row["Kernel Name"])
        n, t = agg.get(k, (0, 0.0))
        agg[k] = (n + 1, t + ns)
        order.append((k, ns))
    total = sum(t for _, t in agg.values())
    print(f"# {title}\n")
    print(f"{sum(n for n, _ in agg.values())} launches, sum {total / 1e6:.2f} ms (cold-cache, serialised under ncu: compare SHARES)\n")
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {n} | {t / 1e3:.1f} | {100 * t / total:.1f}% |")
    groups = collections.Counter()
    for k, (n, t) in agg.items():
        g = ("attention" if "attn" in k else "gemm" if "gemm" in k or "splitk" in k else
             "row kernels" if any(s in k for s in ("ln_modulate", "rmsnorm", "quant_rows")) else "other")
        groups[g] += t
    print("\nBy family: " + ", ".join(f"{g} {100 * t / total:.1f} %" for g, t in groups.most_common()))


if __name__ == "__main__":
    main()

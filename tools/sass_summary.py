"""Per-kernel SASS instruction-count summary of libllb200.so (cuobjdump, no GPU needed): which kernels are
built on the Blackwell tensor / TMA path (UTCHMMA / UTCQMMA = tcgen05.mma bf16 / fp8, LDTM / STTM = tcgen05.ld /
st, UTMALDG = cp.async.bulk.tensor), that no legacy HMMA (mma.sync) is left, plus registers / spills / shared
memory from the resource-usage dump.  Writes a Markdown table (default profiles/r02_sass_summary.md)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "longlive_b200", "libllb200.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "HMMA", "MUFU.EX2",
             "FFMA2", "FMNMX3", "STL", "LDL", "REDUX", "ATOM", "RED"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("llb::", "")


def main():
    out_path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_summary.md")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts, cur, arch = collections.OrderedDict(), None, set()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s*arch = (\S+)", line)
        if m:
            arch.add(m.group(1))
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m:
            op = m.group(1)
            counts[cur]["_total"] += 1
            for mn in MNEMONICS:
                if op == mn or op.startswith(mn + "."):
                    counts[cur][mn] += 1
    res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
    usage, fn = {}, None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and fn:
            usage[fn] = tuple(int(x) for x in m.groups())
    dm = demangle(list(counts))
    cols = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "SYNCS", "MUFU.EX2", "HMMA", "STL", "LDL"]
    lines = ["# SASS summary of longlive_b200/libllb200.so", "",
             f"`cuobjdump -sass` / `--dump-resource-usage`; cubin architectures: {', '.join(sorted(arch))}.",
             "UTCHMMA / UTCQMMA = tcgen05.mma (bf16 / fp8), LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor load,",
             "SYNCS = mbarrier ops, HMMA = legacy mma.sync (must be 0), STL / LDL = local-memory (spill) stores / loads.", "",
             "| kernel | instr | " + " | ".join(cols) + " | regs | stack B | static smem B |",
             "|---|---:|" + "---:|" * (len(cols) + 3)]
    tot = collections.Counter()
    for k, c in counts.items():
        u = usage.get(k, ("?", "?", "?", "?"))
        lines.append(f"| `{short(dm[k])}` | {c['_total']} | " + " | ".join(str(c[m]) for m in cols) +
                     f" | {u[0]} | {u[1]} | {u[2]} |")
        tot.update(c)
    lines += ["", "Totals: " + ", ".join(f"{m} {tot[m]}" for m in MNEMONICS) + f"; kernels {len(counts)}.", ""]
    open(out_path, "w").write("\n".join(lines))
    print("\n".join(lines[-3:]))
    print("wrote", out_path)


if __name__ == "__main__":
    main()

"""Join an ncu source-page capture (SASS level, warp-stall samples per instruction) with nvdisasm's line table, so that the
samples of a kernel can be read per CUDA source line and per code region.

    python tools/ncu_source_lines.py gpurun_out/x.ncu-rep --kernel attn_fwd_kernelILi4ELb1 [--cubin attn_fwd] [--top 40]
        [--regions 'softmax wait S_b:580-590,softmax half:600-700']

Needs the library built with -lineinfo (it is) and the same build that was profiled.  Lines of inlined helpers
(llb_common.cuh) are attributed to the OUTERMOST call site in the .cu file, the helper's own line is kept as a tag.
Runs in the build container (ncu, cuobjdump and nvdisasm are there); nothing here touches the GPU.
"""
import argparse
import collections
import csv
import io
import os
import re
import subprocess
import tempfile

STALLS = ["stall_selected", "stall_wait", "stall_long_sb", "stall_short_sb", "stall_math", "stall_mio", "stall_not_selected",
          "stall_dispatch", "stall_no_inst", "stall_barrier", "stall_membar", "stall_sleep", "stall_branch_resolving",
          "stall_lg", "stall_tex", "stall_drain", "stall_misc"]


def line_table(lib, cubin_key, kernel_key):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
    cub = [f for f in os.listdir(tmp) if cubin_key in f and f.endswith(".cubin")]
    assert cub, f"no cubin matching {cubin_key}"
    sass = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cub[0])], check=True, capture_output=True,
                          text=True).stdout
    table = {}  # offset -> (outer_file, outer_line, inner_tag)
    infn = False
    cur = None
    for ln in sass.splitlines():
        if ln.startswith(".text."):
            infn = kernel_key in ln
            continue
        if not infn:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
        if m:
            inner = (os.path.basename(m.group(1)), int(m.group(2)))
            chain = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            outer = (os.path.basename(chain[-1][0]), int(chain[-1][1])) if chain else inner
            cur = (outer[0], outer[1], f"{inner[0]}:{inner[1]}" if chain else "")
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
        if m and cur is not None:
            table[int(m.group(1), 16)] = cur + (m.group(2).strip(),)
    return table


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--kernel", required=True, help="substring of the mangled kernel name")
    ap.add_argument("--cubin", default="attn_fwd")
    ap.add_argument("--lib", default="longlive_b200/libllb200.so")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--regions", default="", help="name:lo-hi[+lo-hi],... line ranges of the .cu file")
    args = ap.parse_args()
    table = line_table(args.lib, args.cubin, args.kernel)
    out = subprocess.run(["ncu", "-i", args.report, "--page", "source", "--csv"], check=True, capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hi]
    col = {h: i for i, h in enumerate(hdr)}
    base = None
    per_line = collections.defaultdict(lambda: collections.Counter())
    total = 0
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or not r[0]:
            continue
        addr = int(r[0], 16) if r[0].startswith("0x") else int(r[0])
        if base is None:
            base = addr
        off = addr - base
        ent = table.get(off)
        key = (ent[0], ent[1]) if ent else ("?", off)
        n = int(float(r[col["# Samples"]] or 0))
        per_line[key]["samples"] += n
        per_line[key]["inst"] += int(float(r[col["Instructions Executed"]] or 0))
        for s in STALLS:
            if s in col and r[col[s]]:
                per_line[key][s] += int(float(r[col[s]]))
        total += n
    print(f"total samples {total}; {len(per_line)} source lines")
    if args.regions:
        print("\nregions:")
        for spec in args.regions.split(","):
            name, rng = spec.split(":")
            c = collections.Counter()
            for part in rng.split("+"):
                lo, hi_ = (int(x) for x in part.split("-"))
                for (f, l), v in per_line.items():
                    if f.endswith(".cu") and lo <= l <= hi_:
                        c.update(v)
            top = ", ".join(f"{s[6:]} {100 * c[s] / max(c['samples'], 1):.0f}%" for s in sorted(STALLS, key=lambda s: -c[s])[:5])
            print(f"  {name:34s} samples {c['samples']:6d} ({100 * c['samples'] / total:5.1f}%)  warp-inst {c['inst']:9d}  {top}")
    print("\ntop lines:")
    for (f, l), v in sorted(per_line.items(), key=lambda kv: -kv[1]["samples"])[:args.top]:
        top = ", ".join(f"{s[6:]} {100 * v[s] / max(v['samples'], 1):.0f}%" for s in sorted(STALLS, key=lambda s: -v[s])[:3])
        print(f"  {f}:{l:<5} samples {v['samples']:6d} ({100 * v['samples'] / total:5.1f}%) inst {v['inst']:9d}  {top}")


if __name__ == "__main__":
    main()

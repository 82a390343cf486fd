"""Timing experiments on llb_attn_fwd: which resource bounds the steady-state self-attention launch?  attn_fwd.cu
has compile-time knobs (LLB_ATTN_DBG bit mask: parts of the kernel removed, results garbage); this tool builds one
library per mask HERE (`--build`, nvcc cross-compiles) into longlive_b200/_dbg/ and times each on the GPU box
in its own process (`LLB200_LIB` selects the library).

    python tools/attn_debug_sweep.py --build          # in the build container
    gpurun -- python tools/attn_debug_sweep.py        # on the B200
"""
import json
import os
import subprocess
import sys

CHILD = r'''
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(sys.argv[0]))) if False else os.getcwd())
from longlive_b200 import ops
H, Lq, Lk = 12, 4680, 18720
dev = "cuda"
q = torch.randn(Lq, H * 128, device=dev, dtype=torch.bfloat16)
kvs = [(torch.randn(Lk, H * 128, device=dev, dtype=torch.bfloat16), torch.randn(Lk, H * 128, device=dev, dtype=torch.bfloat16)) for _ in range(4)]
out = torch.empty_like(q)
sp = ops.step_params_tensor(ops.make_step_params(attn_segs=[(0, Lk)]), dev)
for i in range(8):
    ops.attention(q, kvs[i % 4][0], kvs[i % 4][1], sp, n_heads=H, out=out)
torch.cuda.synchronize()
st, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
st.record()
for i in range(40):
    ops.attention(q, kvs[i % 4][0], kvs[i % 4][1], sp, n_heads=H, out=out)
en.record(); torch.cuda.synchronize()
print(json.dumps({"lib": os.path.basename(os.environ.get("LLB200_LIB", "libllb200.so")), "ms": st.elapsed_time(en) / 40}))
'''

MODES = (0, 1, 2, 4, 6, 8, 16, 24, 30, 30 + 32, 30 + 64, 30 + 96, 30 + 96 + 128, 31 + 96 + 128, 256, 512, 256 + 24, 512 + 255)
RUN = tuple(int(x) for x in os.environ.get('LLB_DBG_MODES', '').split(',') if x) or MODES
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "longlive_b200", "csrc")
DBG = os.path.join(ROOT, "longlive_b200", "_dbg")  # *.so is git-ignored but travels with gpurun


def build():
    os.makedirs(DBG, exist_ok=True)
    objs = [os.path.join(CSRC, "build", f) for f in os.listdir(os.path.join(CSRC, "build"))
            if f.endswith(".o") and f != "attn_fwd.o"]
    for m in RUN:
        if m == 0 or os.path.exists(os.path.join(DBG, f"libllb200_dbg{m}.so")) and "--force" not in sys.argv:
            continue
        obj = os.path.join(DBG, f"attn_fwd_{m}.o")
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
                               "-Xcompiler", "-fPIC", f"-DLLB_ATTN_DBG={m & 1023}"] + (["-DLLB_MBAR_SPIN"] if m & 1024 else []) +
                              ["-c", os.path.join(CSRC, "attn_fwd.cu"),
                               "-o", obj])
        subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o",
                               os.path.join(DBG, f"libllb200_dbg{m}.so"), obj] + objs)
        os.remove(obj)
        print("built", m, flush=True)


if "--build" in sys.argv:
    build()
    sys.exit(0)

NAMES = {0: "baseline", 1: "no K/V TMA", 2: "no QK MMA", 4: "no PV MMA", 6: "no MMA at all", 8: "no exp2", 16: "no row max",
         24: "no exp2, no max", 9: "no TMA, no exp2", 25: "no TMA, no exp2, no max", 7: "no TMA, no MMA", 30: "no MMA, no exp2, no max"}
res = []
NAMES.update({62: "no MMA, exp2, max, S read", 94: "no MMA, exp2, max, P write", 126: "no MMA, exp2, max, S read, P write",
              254: "barrier skeleton + TMA (no MMA, no softmax work, no TMEM traffic)", 255: "barrier skeleton only",
              256: "same MMA count, N = 16 (1/8 of the tensor work)", 512: "odone committed once per segment",
              280: "N = 16 MMAs, no exp2, no max", 767: "barrier skeleton, odone once per segment",
              1024: "mbarrier test_wait spin instead of try_wait", 1279: "barrier skeleton with test_wait spin"})
for mode in RUN + (0,):
    env = dict(os.environ)
    if mode:
        env["LLB200_LIB"] = os.path.join(DBG, f"libllb200_dbg{mode}.so")
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=300)
    line = [l for l in out.stdout.splitlines() if l.startswith("{")]
    r = json.loads(line[-1]) if line else {"error": out.stderr[-300:]}
    r["debug"] = mode
    r["what"] = NAMES.get(mode, "")
    res.append(r)
    print(r, flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/attn_debug_sweep.json", "w"), indent=1)

"""Hand-off timeline of llb_attn_fwd (instrumented build LLB_ATTN_DBG=2048, see attn_fwd.cu): per key tile, when CTA 0's
softmax warpgroups see S, hand P over, and when the MMA warp sees P and has issued the next MMAs (clock64 of that SM).

    python tools/attn_timeline.py --build      # build container
    gpurun -- python tools/attn_timeline.py    # B200
"""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "longlive_b200", "csrc")
DBG = os.path.join(ROOT, "longlive_b200", "_dbg")
LIB = os.path.join(DBG, "libllb200_dbg2048.so")

if "--build" in sys.argv:
    os.makedirs(DBG, exist_ok=True)
    objs = [os.path.join(CSRC, "build", f) for f in os.listdir(os.path.join(CSRC, "build")) if f.endswith(".o") and f != "attn_fwd.o"]
    obj = os.path.join(DBG, "attn_fwd_2048.o")
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo", "-Xcompiler",
                           "-fPIC", "-DLLB_ATTN_DBG=2048", "-c", os.path.join(CSRC, "attn_fwd.cu"), "-o", obj])
    subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, obj] + objs)
    os.remove(obj)
    sys.exit(0)

os.environ["LLB200_LIB"] = LIB
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from longlive_b200 import _lib, ops  # noqa: E402

H, Lq, Lk = 12, 4680, 18720
dev = "cuda"
q = torch.randn(Lq, H * 128, device=dev, dtype=torch.bfloat16)
k = torch.randn(Lk, H * 128, device=dev, dtype=torch.bfloat16)
v = torch.randn(Lk, H * 128, device=dev, dtype=torch.bfloat16)
out = torch.empty_like(q)
sp = ops.step_params_tensor(ops.make_step_params(attn_segs=[(0, Lk)]), dev)
for _ in range(3):
    ops.attention(q, k, v, sp, n_heads=H, out=out)
torch.cuda.synchronize()
lib = _lib.lib()
buf = (C.c_longlong * (12 * 512))()
lib.llb_attn_debug_ts.restype = C.c_int
n = lib.llb_attn_debug_ts(buf)
ts = [[buf[s * 512 + i] for i in range(512)] for s in range(12)]
names = ["wg0_S_seen", "wg0_P_handed", "mma_P0_seen", "mma_0_issued", "mma_P1_seen", "mma_1_issued", "wg1_S_seen", "wg1_P_handed"]
t0 = ts[0][20]
rows = []
for i in range(20, 60):
    rows.append({nm: ts[s][i] - t0 for s, nm in enumerate(names)})
def avg(f):
    xs = [f(i) for i in range(30, 140)]
    return sum(xs) / len(xs)
summary = {
    "period_per_key_tile": avg(lambda i: ts[0][i + 1] - ts[0][i]),
    "wg0_softmax (S seen -> P handed)": avg(lambda i: ts[1][i] - ts[0][i]),
    "P0 handed -> MMA warp sees it": avg(lambda i: ts[2][i] - ts[1][i]),
    "MMA warp issue PV0 + QK0 (+commits)": avg(lambda i: ts[3][i] - ts[2][i]),
    "MMA issued 0 -> WG0 sees next S (tensor execution + commit + wake)": avg(lambda i: ts[0][i + 1] - ts[3][i]),
    "  wg0: S seen -> S in registers (tcgen05.ld + wait)": avg(lambda i: ts[8][i] - ts[0][i]),
    "  wg0: -> row max / rescale decision": avg(lambda i: ts[9][i] - ts[8][i]),
    "  wg0: -> exp2, sums, bf16 pack, P stores issued": avg(lambda i: ts[10][i] - ts[9][i]),
    "  wg0: -> tcgen05.wait::st": avg(lambda i: ts[11][i] - ts[10][i]),
    "  wg0: -> fence, syncwarp, arrive": avg(lambda i: ts[1][i] - ts[11][i]),
    "wg1_softmax": avg(lambda i: ts[7][i] - ts[6][i]),
    "P1 handed -> MMA warp sees it": avg(lambda i: ts[4][i] - ts[7][i]),
    "MMA warp issue PV1 + QK1 (+commits)": avg(lambda i: ts[5][i] - ts[4][i]),
    "MMA issued 1 -> WG1 sees next S": avg(lambda i: ts[6][i + 1] - ts[5][i]),
    "MMA warp: done with tile 0 -> sees P1 (idle wait)": avg(lambda i: ts[4][i] - ts[3][i]),
    "MMA warp: done with tile 1 -> sees next P0 (idle wait)": avg(lambda i: ts[2][i + 1] - ts[5][i]),
}
for kk, vv in summary.items():
    print(f"{kk:70s} {vv:8.0f} clk")
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"summary_clk": summary, "steps_20_59_relative_clk": rows}, open("gpurun_out/attn_timeline.json", "w"), indent=1)

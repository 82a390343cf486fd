"""The C-ABI shared library loads without a GPU, exports every symbol include/llb200.h declares,
and its structs have the layout the ctypes mirror assumes.  No compute entry point is called."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "llb200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(llb_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from longlive_b200 import _lib
    lib = _lib.lib()
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/llb200.h but not exported"
    assert set(names) == set(_lib.EXPORTED_SYMBOLS), "ctypes prototypes out of sync with the header"
    assert lib.llb_version() == 100


def test_struct_layouts_match_header(tmp_path):
    from longlive_b200 import _lib
    prog = tmp_path / "sz.c"
    prog.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "llb200.h"\n'
        'int main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(llb_kv_state), sizeof(llb_kv_config),'
        ' sizeof(llb_kv_plan), sizeof(llb_step_params), offsetof(llb_kv_plan, n_attn_segs),'
        ' offsetof(llb_step_params, attn_start)); return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(prog), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).split()
    got = [int(x) for x in out]
    exp = [ctypes.sizeof(_lib.KvState), ctypes.sizeof(_lib.KvConfig), ctypes.sizeof(_lib.KvPlan),
           ctypes.sizeof(_lib.StepParams), _lib.KvPlan.n_attn_segs.offset, _lib.StepParams.attn_start.offset]
    assert got == exp


def test_compute_entry_points_fail_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from longlive_b200 import ops
    a = torch.zeros(8, 64, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="CUDA"):
        ops.gemm(a, a)
    from longlive_b200.model import CausalWanModel
    m = CausalWanModel(num_layers=1, dim=256, num_heads=2, ffn_dim=256, text_dim=32, text_len=8)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(torch.zeros(1, 16, 1, 4, 4), t=torch.zeros(1, 1), context=torch.zeros(1, 8, 32),
          kv_cache=[{}], crossattn_cache=[{}], current_start=0)

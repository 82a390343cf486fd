"""The fp32 (cos, sin) table the CUDA RoPE reads == the reference's complex128 `freqs`
(wan/modules/causal_model.py:622-629 built from rope_params, wan/modules/model.py:29-36), rounded
once to fp32.  tests/test_rowkernels_gpu.py builds its fp64 reference rotation from this table, so
the table itself is pinned here: against the oracle's restatement always, against the live reference
module when /root/reference exists (build container)."""
import pytest
import torch

from longlive_b200 import ops
from oracle import ref_shims
from oracle import wan_oracle as wo


def _as_cos_sin(freqs_c128):
    return torch.stack([freqs_c128.real, freqs_c128.imag], dim=-1).to(torch.float32)


@pytest.mark.parametrize("head_dim", [128, 64])
def test_rope_table_equals_oracle_freqs(head_dim):
    cfg = wo.WanConfig(dim=head_dim * 2, num_heads=2)
    ours = ops.build_rope_table(head_dim)
    ref = _as_cos_sin(wo.rope_table(cfg, "cpu"))
    assert ours.shape == ref.shape == (1024, head_dim // 2, 2)
    assert torch.equal(ours, ref)
    # group boundaries: [frame | h | w] = [c - 2*(c//3), c//3, c//3] complex pairs
    c = head_dim // 2
    assert float(ours[0, :, 0].min()) == 1.0 and float(ours[0, :, 1].abs().max()) == 0.0
    for col in (0, c - 2 * (c // 3), c - (c // 3)):  # first pair of every group has frequency 1
        assert torch.allclose(ours[:, col, 0].double(), torch.cos(torch.arange(1024, dtype=torch.float64)), atol=1e-7)


@pytest.mark.skipif(not ref_shims.available(), reason="reference tree not present (GPU box)")
def test_rope_table_equals_live_reference_rope_params():
    _, mm = ref_shims.install()
    d = 128
    freqs = torch.cat([mm.rope_params(1024, d - 4 * (d // 6)), mm.rope_params(1024, 2 * (d // 6)),
                       mm.rope_params(1024, 2 * (d // 6))], dim=1)
    assert freqs.dtype == torch.complex128
    assert torch.equal(ops.build_rope_table(d), _as_cos_sin(freqs))
    assert torch.equal(wo.rope_table(wo.WanConfig(), "cpu"), freqs)

"""Pins oracle/t5_oracle.py (the umT5 text-encoder restatement) to the reference module: the committed fixture
tests/golden/t5_small.pt was produced by the reference's own T5Encoder (oracle/make_t5_golden.py)."""
import os

import pytest
import torch

from oracle import ref_shims
from oracle import t5_oracle as to

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "t5_small.pt")


@pytest.fixture(scope="module")
def golden():
    return torch.load(GOLDEN, weights_only=False)


def test_bucket_table_matches_reference(golden):
    """Integer contract: bucket(key - query) for every pair at text length 512 (t5.py:249-268)."""
    ref = golden["buckets_512"].long()
    tab = to.bucket_table(512)
    i = torch.arange(512)
    mine = tab[(i.unsqueeze(0) - i.unsqueeze(1)) + 511]
    assert torch.equal(mine, ref)
    assert int(tab.min()) == 0 and int(tab.max()) == 31
    assert int(tab[511]) == 0 and int(tab[512]) == 17 and int(tab[510]) == 1


@pytest.mark.parametrize("name,dtype", [("f32", torch.float32), ("bf16", torch.bfloat16)])
def test_oracle_is_bit_identical_to_reference_fixture(golden, name, dtype):
    cfg = to.T5Config(**golden["cfg"])
    sd = to.init_state_dict(cfg, seed=golden["seed"], dtype=dtype, **golden["gains"])
    orc = to.T5EncoderOracle(cfg, sd)
    for (seed, n, b), ref in zip(golden["cases"], golden[name]):
        ids, mask = to.synth_token_ids(cfg, seed, n, b)
        out = orc.text_encoder_forward(ids, mask)["prompt_embeds"]
        assert out.dtype == dtype and out.shape == ref.shape
        assert torch.equal(out, ref), float((out.float() - ref.float()).abs().max())
        lens = mask.sum(1)
        for r in range(b):
            assert float(out[r, lens[r]:].abs().max() if lens[r] < cfg.text_len else 0.0) == 0.0
            assert float(out[r, :lens[r]].abs().min()) >= 0.0


def test_fixture_attention_is_not_degenerate(golden):
    """The fixture's gains must give a peaked softmax and a position bias that matters, otherwise the
    attention path would not be pinned by it."""
    cfg = to.T5Config(**golden["cfg"])
    sd = to.init_state_dict(cfg, seed=golden["seed"], dtype=torch.float32, **golden["gains"])
    ids, mask = to.synth_token_ids(cfg, 12, 96, 1)
    x = torch.nn.functional.embedding(ids, sd["token_embedding.weight"])
    xn = to.t5_layer_norm(x, sd["blocks.0.norm1.weight"], cfg.eps)
    q = (xn @ sd["blocks.0.attn.q.weight"].T).view(1, -1, cfg.num_heads, cfg.head_dim)
    k = (xn @ sd["blocks.0.attn.k.weight"].T).view(1, -1, cfg.num_heads, cfg.head_dim)
    logits = torch.einsum("binc,bjnc->bnij", q, k)
    assert float(logits.std()) > 1.0
    assert float(sd["blocks.0.pos_embedding.embedding.weight"].std()) > 0.2
    # padding does not leak: rows below seq_len are the same whether or not other rows are padding
    orc = to.T5EncoderOracle(cfg, sd)
    ids2, mask2 = to.synth_token_ids(cfg, 12, 40, 1)
    a = orc.encode(ids2, mask2)[:, :40]
    ids3 = ids2.clone(); ids3[:, 40:] = 5
    b = orc.encode(ids3, mask2)[:, :40]
    assert torch.equal(a, b)


@pytest.mark.skipif(not ref_shims.available(), reason="reference tree not present (GPU box)")
def test_oracle_matches_live_reference():
    from oracle import make_t5_golden as mk
    mod = mk.load_reference_t5()
    cfg = to.T5Config(vocab=500, dim=128, dim_attn=128, dim_ffn=256, num_heads=2, num_layers=2, text_len=40)
    sd = to.init_state_dict(cfg, seed=3, dtype=torch.bfloat16, q_gain=32.0, pos_gain=16.0)
    enc = mk.reference_encoder(mod, cfg, sd, torch.bfloat16)
    ids, mask = to.synth_token_ids(cfg, 5, 23, 2)
    ref = mk.reference_text_encoder_forward(enc, ids, mask)
    out = to.T5EncoderOracle(cfg, sd).text_encoder_forward(ids, mask)["prompt_embeds"]
    assert torch.equal(out, ref)

"""Host-side logic of the umT5 text-encoder mirror (no GPU): integer bucket table, parameter names, tokenizer
contract, and that the compute path refuses to run without CUDA."""
import os

import pytest
import torch

from longlive_b200 import synth
from longlive_b200.text_encoder import HashTokenizer, UMT5Encoder, WanTextEncoder, relative_position_buckets
from oracle import t5_oracle as to

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "t5_small.pt")
SMALL = dict(vocab=300, dim=128, dim_attn=128, dim_ffn=256, num_heads=2, num_layers=2, text_len=40)


def test_bucket_table_bit_exact():
    gold = torch.load(GOLDEN, weights_only=False)["buckets_512"].long()
    lut = relative_position_buckets(512).long()
    assert lut.dtype == torch.int64 and lut.numel() == 1023
    i = torch.arange(512)
    assert torch.equal(lut[(i.unsqueeze(0) - i.unsqueeze(1)) + 511], gold)      # reference module's own output
    assert torch.equal(lut, to.bucket_table(512))                                # oracle restatement
    wide = relative_position_buckets(640).long()
    assert torch.equal(wide[640 - 512:640 + 511], lut)                           # longer table = same offsets


def test_parameter_names_and_shapes_match_reference_state_dict():
    cfg = to.T5Config(**SMALL)
    sd = to.init_state_dict(cfg, seed=0)     # keys / shapes verified against the reference module by load_state_dict(strict=True)
    enc = UMT5Encoder(**SMALL)
    mine = enc.state_dict()
    assert set(mine) == set(sd)
    for k in sd:
        assert tuple(mine[k].shape) == tuple(sd[k].shape), k
    enc.load_state_dict(sd, strict=True)
    full = UMT5Encoder.__init__.__defaults__
    assert full[:7] == (256384, 4096, 4096, 10240, 64, 24, 32)   # umt5_xxl (t5.py:459-472)


def test_random_init_follows_reference_distributions():
    enc = UMT5Encoder(**SMALL)
    synth.random_init_t5_(enc, seed=1)
    sd = enc.state_dict()
    assert abs(float(sd["token_embedding.weight"].std()) - 1.0) < 0.05
    assert abs(float(sd["blocks.0.attn.k.weight"].std()) * 128 ** 0.5 - 1.0) < 0.1
    assert abs(float(sd["blocks.1.ffn.fc2.weight"].std()) * 256 ** 0.5 - 1.0) < 0.1
    assert torch.equal(sd["blocks.0.norm1.weight"], torch.ones(128))


def test_hash_tokenizer_contract():
    tok = HashTokenizer(seq_len=16, vocab_size=1000)
    ids, mask = tok(["a b c", "d " * 40], return_mask=True, add_special_tokens=True)
    assert ids.shape == mask.shape == (2, 16) and ids.dtype == torch.long
    assert mask[0].tolist() == [1] * 4 + [0] * 12 and int(ids[0, 3]) == 1 and int(ids[0, 4:].abs().sum()) == 0
    assert int(mask[1].sum()) == 16 and int(ids[1, 15]) == 1
    ids2, _ = tok(["a b c"], return_mask=True)
    assert torch.equal(ids2[0], ids[0])          # deterministic
    assert int(ids.max()) < 1000 and int(ids[mask.bool()].min()) >= 1


def test_encoder_fails_loudly_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    enc = UMT5Encoder(**SMALL).to(torch.bfloat16)
    te = WanTextEncoder(text_encoder=enc, tokenizer=HashTokenizer(seq_len=40, vocab_size=300))
    with pytest.raises(RuntimeError, match="CUDA"):
        te(["a synthetic prompt"])
    with pytest.raises(FileNotFoundError):
        WanTextEncoder()                          # no checkpoint offline: must not silently random-init


def test_non_prefix_mask_rejected():
    enc = UMT5Encoder(**SMALL)
    enc._packed = {"emb": torch.zeros(1)}        # skip packing: the mask check happens before any kernel
    ids = torch.ones(1, 8, dtype=torch.long)
    mask = torch.tensor([[1, 0, 1, 0, 0, 0, 0, 0]])
    with pytest.raises(ValueError, match="prefix"):
        enc(ids, mask)


def test_geglu_weight_layout():
    """EPI_GEGLU_BF16 expects 256-row tiles of [128 gate rows | the 128 fc1 rows of the same output columns]."""
    from longlive_b200 import ops
    F, K = 384, 16
    gate = torch.arange(F * K, dtype=torch.float32).view(F, K)
    fc1 = -gate
    w = ops.geglu_weight(gate, fc1)
    assert w.shape == (2 * F, K)
    for t in range(F // 128):
        assert torch.equal(w[t * 256:t * 256 + 128], gate[t * 128:(t + 1) * 128])
        assert torch.equal(w[t * 256 + 128:(t + 1) * 256], fc1[t * 128:(t + 1) * 128])
    with pytest.raises(AssertionError):
        ops.geglu_weight(gate[:100], fc1[:100])

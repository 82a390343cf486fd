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


class _RecordingOps:
    """Stands in for longlive_b200.ops: records the launch sequence of UMT5Encoder._run without a GPU."""
    EPI_BIAS, EPI_BIAS_RES, EPI_BIAS_GELU_BF16, EPI_BIAS_MUL, EPI_GEGLU_BF16 = 0, 4, 7, 6, 8

    def __init__(self):
        self.calls = []

    def embed_rows(self, table, ids, rows, out=None):
        self.calls.append(("embed",)); return out

    def rmsnorm(self, x, w, eps, out=None):
        self.calls.append(("norm", id(w))); return out

    def gemm(self, a, w, bias=None, *, epilogue=0, out=None, res=None, **kw):
        self.calls.append(("gemm", id(w), epilogue)); return out

    def gemm_splitk(self, a, w, ws, k_splits, bias=None, *, res=None, out=None, norm_w=None, norm_out=None, norm_eps=0.0):
        self.calls.append(("splitk", id(w), k_splits, id(norm_w) if norm_out is not None else None)); return out

    def t5_attention(self, qkv, B, H, lens, pos, lut, out=None):
        self.calls.append(("attn",)); return out

    def t5_final_norm(self, x, w, B, rows_out, lens, eps, out=None):
        self.calls.append(("final", id(w))); return out


@pytest.mark.parametrize("rows,expect_split,expect_geglu", [(128, True, True), (256, True, False), (512, False, True)])
def test_encoder_launch_sequence(monkeypatch, rows, expect_split, expect_geglu):
    """Every block normalises exactly what the reference normalises (t5.py:163-168), whichever fusions are active:
    short prompts use the split-K projections whose reduce launch applies the following norm, the gated FFN is one
    launch except at 129-256 rows."""
    import longlive_b200.text_encoder as te
    rec = _RecordingOps()
    monkeypatch.setattr(te, "ops", rec)
    enc = UMT5Encoder(vocab=50, dim=128, dim_attn=128, dim_ffn=256, num_heads=2, num_layers=3, text_len=512)
    z = lambda *s: torch.zeros(*s)
    layers = [{k: z(1) for k in ("n1", "n2", "qkv", "o", "gate", "fc1", "fc2", "pos", "gf")} for _ in range(3)]
    enc._packed = {"emb": z(1), "norm": z(1), "lut": z(1), "layers": layers}
    b = {k: z(1) for k in ("x", "xn", "qkv", "att", "g", "h", "out", "ids", "lens")}
    b["ws"] = z(1) if rows <= 256 else None
    enc._run(b, 1, rows, 512, True)
    names = [c[0] for c in rec.calls]
    assert names[0] == "embed" and names[-1] == "final" and names.count("attn") == 3
    # norm1 of block 0 is always its own launch; later norm1s and every norm2 are either a launch or fused into a reduce
    norm_launch = [c[1] for c in rec.calls if c[0] == "norm"]
    fused = [c[3] for c in rec.calls if c[0] == "splitk" and c[3] is not None]
    want = []
    for lw in layers:
        want += [id(lw["n1"]), id(lw["n2"])]
    assert sorted(norm_launch + fused) == sorted(want)
    assert norm_launch[0] == id(layers[0]["n1"])
    # order inside a block: (norm1) qkv attn o (norm2) ffn fc2
    i = 0
    for li, lw in enumerate(layers):
        if li == 0 or not expect_split:
            assert rec.calls[i + 1] == ("norm", id(lw["n1"])); i += 1
        assert rec.calls[i + 1][:2] == ("gemm", id(lw["qkv"])); assert rec.calls[i + 2] == ("attn",); i += 2
        if expect_split:
            assert rec.calls[i + 1] == ("splitk", id(lw["o"]), 2, id(lw["n2"])); i += 1
        else:
            assert rec.calls[i + 1] == ("gemm", id(lw["o"]), rec.EPI_BIAS_RES); assert rec.calls[i + 2] == ("norm", id(lw["n2"])); i += 2
        if expect_geglu:
            assert rec.calls[i + 1] == ("gemm", id(lw["gf"]), rec.EPI_GEGLU_BF16); i += 1
        else:
            assert rec.calls[i + 1] == ("gemm", id(lw["gate"]), rec.EPI_BIAS_GELU_BF16)
            assert rec.calls[i + 2] == ("gemm", id(lw["fc1"]), rec.EPI_BIAS_MUL); i += 2
        if expect_split:
            nxt = id(layers[li + 1]["n1"]) if li + 1 < 3 else None
            assert rec.calls[i + 1] == ("splitk", id(lw["fc2"]), 2, nxt); i += 1
        else:
            assert rec.calls[i + 1] == ("gemm", id(lw["fc2"]), rec.EPI_BIAS_RES); i += 1
    assert rec.calls[i + 1] == ("final", id(enc._packed["norm"]))

"""ORACLE support — test infrastructure only.

Imports the REAL reference modules from /root/reference (read-only, present in the build container
only, never on the GPU box) so that oracle/make_golden.py can (a) pin oracle/wan_oracle.py against
them and (b) generate the committed fixtures under tests/golden/.  Nothing is copied: the reference
code runs from where it lies.  The shims only neutralise imports that cannot be satisfied here
(SURVEY.md section 8c): heavy package __init__s, `diffusers` base classes, `utils.memory`'s
import-time CUDA query, `ftfy`, and the CUDA-only flash-attn entry point.
"""
from __future__ import annotations

import importlib
import os
import sys
import types

import torch

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# the unmodified reference tree as shipped to the GPU box: __graft_entry__.build() mirrors
# /root/reference (minus assets) into baseline/_ref/ (git-ignored, travels with gpurun)
SHIPPED_ROOT = os.path.join(_REPO, "baseline", "_ref")


def _has_tree(root: str) -> bool:
    return os.path.isdir(os.path.join(root, "wan", "modules"))


def _default_root() -> str:
    env = os.environ.get("LLB_REFERENCE_ROOT")
    if env:
        return env
    return "/root/reference" if _has_tree("/root/reference") else SHIPPED_ROOT


REFERENCE_ROOT = _default_root()


def available() -> bool:
    return _has_tree(REFERENCE_ROOT)


def shipped_available() -> bool:
    return _has_tree(SHIPPED_ROOT)


def use_shipped_copy() -> None:
    """GPU tests / bench.py: import the reference from baseline/_ref only (never /root/reference)."""
    global REFERENCE_ROOT
    if not shipped_available():
        raise RuntimeError(f"no shipped reference tree at {SHIPPED_ROOT}: run __graft_entry__.build() in the "
                           "build container first")
    REFERENCE_ROOT = SHIPPED_ROOT


def _stub(name: str, **attrs) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def install(attention_impl: str = "sdpa"):
    """Makes `wan.modules.causal_model`, `utils.wan_wrapper`, `pipeline.*` importable on CPU.

    attention_impl: 'sdpa' -> torch SDPA stands in for flash-attn (same maths: scale 1/sqrt(d), no
    mask); 'exact' -> oracle.wan_oracle.exact_attention (used to show that everything *around*
    attention is bit-identical between the oracle and the reference); 'flash' -> nothing is patched:
    the reference's own flash-attn call (wan/modules/attention.py:131-145), CUDA only.
    """
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)

    # (1) package shells so wan/__init__.py and wan/modules/__init__.py do not execute
    for pkg, sub in (("wan", "wan"), ("wan.modules", os.path.join("wan", "modules"))):
        if pkg not in sys.modules:
            m = _stub(pkg)
            m.__path__ = [os.path.join(REFERENCE_ROOT, sub)]
    # (2) diffusers base classes used only as mixins / decorator
    if "diffusers" not in sys.modules:
        class ConfigMixin:  # noqa: D401
            pass

        class ModelMixin(torch.nn.Module):
            pass

        def register_to_config(fn):
            return fn

        _stub("diffusers")
        _stub("diffusers.configuration_utils", ConfigMixin=ConfigMixin, register_to_config=register_to_config)
        _stub("diffusers.models")
        _stub("diffusers.models.modeling_utils", ModelMixin=ModelMixin)
    # (3) utils.memory queries the CUDA device at import time
    if not torch.cuda.is_available() and "utils.memory" not in sys.modules:
        if "utils" not in sys.modules:
            u = _stub("utils")
            u.__path__ = [os.path.join(REFERENCE_ROOT, "utils")]
        noop = lambda *a, **k: None
        _stub("utils.memory", gpu=torch.device("cpu"), get_cuda_free_memory_gb=lambda *a, **k: 0.0,
              DynamicSwapInstaller=object, log_gpu_memory=noop,
              move_model_to_device_with_memory_preservation=noop)
    # (5) ftfy is imported by the tokenizer module
    if "ftfy" not in sys.modules:
        try:
            importlib.import_module("ftfy")
        except Exception:
            _stub("ftfy", fix_text=lambda s: s)

    cm = importlib.import_module("wan.modules.causal_model")
    mm = importlib.import_module("wan.modules.model")

    # (4) attention entry points
    if attention_impl == "flash":
        if not torch.cuda.is_available():
            raise RuntimeError("attention_impl='flash' runs the reference's flash-attn path: CUDA only")
        am = importlib.import_module("wan.modules.attention")
        cm.attention = am.attention                  # undo an earlier patch in this process
        mm.flash_attention = am.flash_attention
        return cm, mm
    if attention_impl == "fallback":
        # the reference's OWN no-flash-attn branch (wan/modules/attention.py:183-197: torch SDPA), selected by
        # clearing its availability flags; the direct flash_attention() call of the cross-attention
        # (wan/modules/model.py:189) asserts CUDA, so on a CPU box it is routed through the same branch
        am = importlib.import_module("wan.modules.attention")
        am.FLASH_ATTN_2_AVAILABLE = am.FLASH_ATTN_3_AVAILABLE = False
        cm.attention = am.attention
        mm.flash_attention = lambda q, k, v, *a, **kw: am.attention(q, k, v)
        return cm, mm
    if attention_impl == "exact":
        from oracle.wan_oracle import exact_attention

        def attn(q, k, v, *a, **kw):
            return exact_attention(q, k, v)
    else:
        def attn(q, k, v, *a, **kw):
            o = torch.nn.functional.scaled_dot_product_attention(
                q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2))
            return o.transpose(1, 2).contiguous()
    cm.attention = attn
    mm.flash_attention = attn
    return cm, mm


def build_reference_model(cfg, state_dict, attention_impl: str = "sdpa", device=None):
    """Instantiates the reference CausalWanModel with the oracle's config + weights (bf16, eval).
    device: construct (and run init_weights) directly on that device - seconds instead of a minute
    for the 1.4 B-parameter shape."""
    import contextlib
    cm, _ = install(attention_impl)
    with (torch.device(device) if device is not None else contextlib.nullcontext()):
        model = _construct(cm, cfg)
    if device is not None:
        state_dict = {k: v.to(device) for k, v in state_dict.items()}
    return _finish(model, cfg, state_dict)


def _construct(cm, cfg):
    return cm.CausalWanModel(
        model_type="t2v", patch_size=cfg.patch, text_len=cfg.text_len, in_dim=cfg.in_dim, dim=cfg.dim,
        ffn_dim=cfg.ffn_dim, freq_dim=cfg.freq_dim, text_dim=cfg.text_dim, out_dim=cfg.out_dim,
        num_heads=cfg.num_heads, num_layers=cfg.num_layers, local_attn_size=cfg.local_attn_size,
        sink_size=cfg.sink_size, qk_norm=True, cross_attn_norm=True, eps=cfg.eps)


def _finish(model, cfg, state_dict):
    missing, unexpected = model.load_state_dict(state_dict, strict=False)
    assert not unexpected, unexpected
    assert all("freqs" in m for m in missing), missing
    model = model.to(torch.bfloat16).eval()
    # the reference hard-codes 1560 tokens per frame in max_attention_size (causal_model.py:88);
    # small-grid fixtures set it the way the pipeline does (_set_all_modules_max_attention_size)
    for mod in model.modules():
        if hasattr(mod, "max_attention_size"):
            mod.max_attention_size = cfg.max_attention_size
    return model


def build_reference_wrapper(cfg, state_dict, shift: float = 5.0, attention_impl: str = "sdpa", device=None):
    """The reference's WanDiffusionWrapper around a directly-constructed model (its __init__ needs
    checkpoint files, utils/wan_wrapper.py:132-133)."""
    model = build_reference_model(cfg, state_dict, attention_impl, device)
    cur = torch.cuda.current_device
    if not torch.cuda.is_available():
        torch.cuda.current_device = lambda: 0  # wan/modules/t5.py evaluates it at class-definition time
    try:
        ww = importlib.import_module("utils.wan_wrapper")
        sch = importlib.import_module("utils.scheduler")
    finally:
        torch.cuda.current_device = cur
    w = ww.WanDiffusionWrapper.__new__(ww.WanDiffusionWrapper)
    torch.nn.Module.__init__(w)
    w.model = model
    w.uniform_timestep = False
    w.scheduler = sch.FlowMatchScheduler(shift=shift, sigma_min=0.0, extra_one_step=True)
    w.scheduler.set_timesteps(1000, training=True)
    w.seq_len = 32760
    w.post_init()
    return w


def reference_pipelines():
    cur = torch.cuda.current_device
    if not torch.cuda.is_available():
        torch.cuda.current_device = lambda: 0
    try:
        ci = importlib.import_module("pipeline.causal_inference")
        ii = importlib.import_module("pipeline.interactive_causal_inference")
    finally:
        torch.cuda.current_device = cur
    return ci.CausalInferencePipeline, ii.InteractiveCausalInferencePipeline

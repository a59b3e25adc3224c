"""Import the REAL reference from /root/reference (build container only; absent on the GPU box).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Used by oracle/make_goldens.py to pin the
restatement and to produce tests/golden/*.  The reference's ``models/__init__.py`` pulls in
diffusers and omegaconf, which are not installed, so the package is registered as a stub and its
sub-modules are imported directly.  Nothing is copied: the reference's
source files are executed where they lie.
"""
from __future__ import annotations

import contextlib
import dataclasses
import importlib
import io
import os
import sys
import types

import torch
import torch.nn as nn

REF = os.environ.get("MMADA_REFERENCE", "/root/reference")


def quiet():
    """Context manager that swallows the reference's prints (config dumps, debugging output)."""
    return contextlib.redirect_stdout(io.StringIO())


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "models", "modeling_mmada.py"))


def _stub_packages():
    if "models" in sys.modules and getattr(sys.modules["models"], "__mmada_stub__", False):
        return
    pkg = types.ModuleType("models")
    pkg.__path__ = [os.path.join(REF, "models")]
    pkg.__mmada_stub__ = True
    sys.modules["models"] = pkg
    mu = types.ModuleType("models.modeling_utils")
    mu.ConfigMixin = type("ConfigMixin", (), {})
    mu.ModelMixin = type("ModelMixin", (nn.Module,), {})
    mu.register_to_config = lambda f: f
    sys.modules["models.modeling_utils"] = mu
    if "omegaconf" not in sys.modules:
        oc = types.ModuleType("omegaconf")
        oc.OmegaConf = object
        oc.DictConfig = dict
        sys.modules["omegaconf"] = oc


def modules():
    """(modeling_mmada, configuration_llada, sampling, modeling_magvitv2) of the reference."""
    _stub_packages()
    mm = importlib.import_module("models.modeling_mmada")
    cl = importlib.import_module("models.configuration_llada")
    sp = importlib.import_module("models.sampling")
    mv = importlib.import_module("models.modeling_magvitv2")
    return mm, cl, sp, mv


def build_model(cfg: dict, state_dict=None):
    """Reference ``MMadaModelLM`` (fp32, CPU, eval) for an oracle config dict, loaded with ``state_dict``."""
    mm, cl, _, _ = modules()
    base = dataclasses.asdict(cl.ModelConfig())
    base.update(d_model=cfg["d_model"], n_heads=cfg["n_heads"], n_kv_heads=cfg["n_heads"],
                n_layers=cfg["n_layers"], mlp_hidden_size=cfg["mlp_hidden_size"],
                vocab_size=cfg["vocab_size"], embedding_size=cfg["vocab_size"], block_type="llama",
                activation_type="silu", layer_norm_type="rms", rope=True, rope_theta=cfg["rope_theta"],
                rms_norm_eps=cfg["rms_norm_eps"], weight_tying=False, include_bias=False,
                max_sequence_length=cfg["max_sequence_length"], mask_token_id=cfg["mask_token_id"],
                attention_dropout=0.0, residual_dropout=0.0, embedding_dropout=0.0, init_device="cpu",
                init_fn="mitchell")
    with contextlib.redirect_stdout(io.StringIO()):
        c = mm.MMadaConfig(**base)
        c.use_cache = False
        model = mm.MMadaModelLM(c, init_params=False).eval()
    if state_dict is not None:
        missing, unexpected = model.load_state_dict(state_dict, strict=False)
        assert not unexpected, unexpected
        assert all("rotary" in m or "inv_freq" in m for m in missing), missing
    return model


def build_vq(state_dict=None):
    _, _, _, mv = modules()
    with contextlib.redirect_stdout(io.StringIO()):
        vq = mv.MAGVITv2().eval()
    if state_dict is not None:
        missing, unexpected = vq.load_state_dict(state_dict, strict=False)
        assert not unexpected, unexpected
        have = {k.split(".")[0] for k in state_dict}
        assert all(m.split(".")[0] not in have or m.startswith("quantize.") for m in missing), missing
    return vq


def load_generate_fn():
    """The reference's free function ``generate`` (generate.py:43), with its package import removed."""
    _stub_packages()
    src = open(os.path.join(REF, "generate.py")).read().replace("from models import MMadaModelLM", "")
    ns: dict = {}
    exec(compile(src, os.path.join(REF, "generate.py"), "exec"), ns)
    return ns["generate"]


class UniPromptingStub:
    """All ``t2i_generate`` reads from ``uni_prompting``: ``len(text_tokenizer)`` (modeling_mmada.py:149)."""

    class _Tok:
        def __init__(self, n):
            self.n = n

        def __len__(self):
            return self.n

    def __init__(self, text_vocab: int = 126349):
        self.text_tokenizer = self._Tok(text_vocab)

"""Synthetic model configurations and weights shared by the oracle, the goldens and the tests.

TEST INFRASTRUCTURE (see oracle/__init__.py).  There is no network, hence no
checkpoint: every parity case runs on random weights.  The generator below is
deliberately independent of the reference's own ``init_weights``
(/root/reference/models/modeling_llada.py:80-155) so that it can be re-run on the
GPU box, where the reference is absent, and give bit-identical tensors: each
tensor gets its own ``torch.Generator`` seeded from (seed, key name).  The scale
follows the reference's "mitchell" rule (std = 1/sqrt(fan_in), residual outputs
further divided by sqrt(2*(layer+1)), truncation at 3 std) so logits have a
realistic O(1) scale.

Key names are the reference's state-dict names so the same
dict loads into the real ``MMadaModelLM`` / ``MAGVITv2`` with ``load_state_dict``.
"""
from __future__ import annotations

import hashlib
import math
from collections.abc import Mapping
from typing import Dict

import torch

# --- configurations -----------------------------------------------------------------------

#: BASELINE.json configs[0]: reduced LLaDA/MMaDA, 4 layers, d=1024 (heads / ffn pinned here).
C1 = dict(d_model=1024, n_heads=16, n_layers=4, mlp_hidden_size=2816, vocab_size=134656,
          rope_theta=500000.0, rms_norm_eps=1e-5, max_sequence_length=4096, mask_token_id=126336)

#: BASELINE.json configs[1]: MMaDA-8B architecture (configs/mmada_demo.yaml + LLaDA-8B config).
C2 = dict(d_model=4096, n_heads=32, n_layers=32, mlp_hidden_size=12288, vocab_size=134656,
          rope_theta=500000.0, rms_norm_eps=1e-5, max_sequence_length=4096, mask_token_id=126336)

#: tiny model for fast unit tests (head_dim 64) and its head_dim-128 sibling
TINY = dict(d_model=256, n_heads=4, n_layers=2, mlp_hidden_size=512, vocab_size=134656,
            rope_theta=500000.0, rms_norm_eps=1e-5, max_sequence_length=4096, mask_token_id=126336)
TINY128 = dict(TINY, d_model=512, n_heads=4, mlp_hidden_size=1024)
#: text-to-motion: the vocabulary must hold text + 8192 image codes + 512 motion codes
TINY_T2M = dict(TINY, vocab_size=135168)

#: `len(uni_prompting.text_tokenizer)` stand-in (reference app.py:396); image code c has id c + this.
TEXT_VOCAB = 126349
CODEBOOK = 8192
# reserved token ids (reference training/prompting_utils.py:17-33)
TOK_SOI, TOK_EOI, TOK_T2I, TOK_MMU, TOK_PAD = 126084, 126085, 126088, 126089, 126093
TOK_BOS, TOK_EOS = 126080, 126081


def _gen(seed: int, name: str) -> torch.Generator:
    h = hashlib.sha256(f"{seed}:{name}".encode()).digest()
    g = torch.Generator(device="cpu")
    g.manual_seed(int.from_bytes(h[:7], "little"))
    return g


def _normal(seed: int, name: str, shape, std: float, dtype=torch.float32, device="cpu") -> torch.Tensor:
    """Truncated normal from the (seed, name) generator.  ``device`` other than the CPU draws on that device from a
    generator with the same seed: same rule, different (device-specific) values — only for cases without goldens."""
    device = torch.device(device)
    if device.type == "cpu":
        t = torch.randn(shape, generator=_gen(seed, name), dtype=torch.float32)
    else:
        g = torch.Generator(device=device)
        g.manual_seed(_gen(seed, name).initial_seed())
        t = torch.randn(shape, generator=g, dtype=torch.float32, device=device)
    t.clamp_(-3.0, 3.0).mul_(std)
    return t.to(dtype)


def llada_weight_keys(cfg: dict):
    """State-dict key names of the reference model for ``cfg``, in load order."""
    p = "model.transformer."
    keys = [p + "wte.weight"]
    for i in range(cfg["n_layers"]):
        b = f"{p}blocks.{i}."
        keys += [b + n + ".weight" for n in ("attn_norm", "ff_norm", "q_proj", "k_proj", "v_proj", "attn_out", "ff_proj",
                                             "up_proj", "ff_out")]
    return keys + [p + "ln_f.weight", p + "ff_out.weight"]


def llada_weight(cfg: dict, seed: int, key: str, dtype=torch.float32, device="cpu") -> torch.Tensor:
    """ONE tensor of ``make_llada_weights(cfg, seed)`` by its state-dict key (same values on the CPU)."""
    d, ffn, V = cfg["d_model"], cfg["mlp_hidden_size"], cfg["vocab_size"]
    p = "model.transformer."
    if key == p + "wte.weight":
        return _normal(seed, "wte", (V, d), 1.0 / math.sqrt(d), dtype, device)
    if key == p + "ln_f.weight":
        return (1.0 + _normal(seed, "ln_f", (d,), 0.1, device=device)).to(dtype)
    if key == p + "ff_out.weight":
        return _normal(seed, "head", (V, d), 1.0 / math.sqrt(d), dtype, device)
    assert key.startswith(p + "blocks.") and key.endswith(".weight"), key
    i, n = key[len(p + "blocks."):-len(".weight")].split(".")
    i = int(i)
    b = f"{p}blocks.{i}."
    res = 1.0 / math.sqrt(2 * (i + 1))
    if n in ("attn_norm", "ff_norm"):
        return (1.0 + _normal(seed, b + n, (d,), 0.1, device=device)).to(dtype)
    shape, std = {"q_proj": ((d, d), 1.0 / math.sqrt(d)), "k_proj": ((d, d), 1.0 / math.sqrt(d)),
                  "v_proj": ((d, d), 1.0 / math.sqrt(d)), "attn_out": ((d, d), res / math.sqrt(d)),
                  "ff_proj": ((ffn, d), 1.0 / math.sqrt(d)), "up_proj": ((ffn, d), 1.0 / math.sqrt(d)),
                  "ff_out": ((d, ffn), res / math.sqrt(ffn))}[n]
    return _normal(seed, b + n, shape, std, dtype, device)


class LazyLladaWeights(Mapping):
    """``make_llada_weights`` as a lazy mapping: a tensor is generated when it is read and not kept, so an 8B-parameter
    state dict (32 GB in fp32) can be streamed layer by layer (tests/test_full_size_gpu.py)."""

    def __init__(self, cfg: dict, seed: int = 0, dtype=torch.float32, device="cpu"):
        self.cfg, self.seed, self.dtype, self.device = cfg, seed, dtype, device
        self._keys = llada_weight_keys(cfg)

    def __getitem__(self, key: str) -> torch.Tensor:
        return llada_weight(self.cfg, self.seed, key, self.dtype, self.device)

    def __iter__(self):
        return iter(self._keys)

    def __len__(self):
        return len(self._keys)

    def block(self, i: int) -> Dict[str, torch.Tensor]:
        """The nine tensors of block ``i``, materialised."""
        b = f"model.transformer.blocks.{i}."
        return {k: self[k] for k in self._keys if k.startswith(b)}


def make_llada_weights(cfg: dict, seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """State dict with the reference's key names for ``cfg`` (no biases, untied head)."""
    return {k: llada_weight(cfg, seed, k, dtype) for k in llada_weight_keys(cfg)}


def make_t2i_prompts(batch: int, prefix_len: int, n_img: int, seed: int = 0, mask_id: int = 126336):
    """Synthetic t2i_gen rows laid out like reference training/prompting_utils.py:200-233:
    ``[pad.. <|t2i|> bos text eos]`` (prefix_len ids, left padded) ``<|soi|> mask*n_img <|eoi|>``.
    Returns (cond_ids, uncond_ids, cond_attn, uncond_attn), all int64 (batch, prefix_len+n_img+2)."""
    g = _gen(seed, "prompts")
    L = prefix_len + 1 + n_img + 1
    cond = torch.full((batch, L), TOK_PAD, dtype=torch.int64)
    unc = torch.full((batch, L), TOK_PAD, dtype=torch.int64)
    for b in range(batch):
        t = int(torch.randint(4, min(64, prefix_len - 3) + 1, (1,), generator=g))
        text = torch.randint(0, 126000, (t,), generator=g)
        row = torch.cat([torch.tensor([TOK_T2I, TOK_BOS]), text, torch.tensor([TOK_EOS])])
        cond[b, prefix_len - row.numel():prefix_len] = row
        urow = torch.tensor([TOK_T2I, TOK_BOS, TOK_EOS])
        unc[b, prefix_len - 3:prefix_len] = urow
    for ids in (cond, unc):
        ids[:, prefix_len] = TOK_SOI
        ids[:, prefix_len + 1:prefix_len + 1 + n_img] = mask_id
        ids[:, -1] = TOK_EOI
    return cond, unc, (cond != TOK_PAD).long(), (unc != TOK_PAD).long()


# --- MAGVIT-v2 decoder --------------------------------------------------------------------

VQ_CH, VQ_CH_MULT, VQ_NUM_RES, VQ_Z = 128, (1, 1, 2, 2, 4), (4, 4, 3, 4, 3), 13


def vq_decoder_plan():
    """Decoder topology as a flat list of (kind, key-prefix, c_in, c_out), following
    reference models/modeling_magvitv2.py:309-362 / :365-399 (up levels iterated 4..0)."""
    plan = [("conv1", "post_quant_conv", VQ_Z, VQ_Z)]
    c = VQ_CH * VQ_CH_MULT[-1]
    plan += [("conv3", "conv_in", VQ_Z, c), ("res", "mid.block_1", c, c), ("attn", "mid.attn_1", c, c),
             ("res", "mid.block_2", c, c)]
    for lvl in reversed(range(5)):
        co = VQ_CH * VQ_CH_MULT[lvl]
        for j in range(VQ_NUM_RES[lvl]):
            plan.append(("res", f"up.{lvl}.block.{j}", c, co))
            c = co
        if lvl != 0:
            plan.append(("up", f"up.{lvl}.upsample", c, c))
    plan += [("norm_out", "norm_out", c, c), ("conv3", "conv_out", c, 3)]
    return plan


def make_vq_decoder_weights(seed: int = 0) -> Dict[str, torch.Tensor]:
    """fp32 state dict for ``MAGVITv2.decoder`` with the reference's key names, prefixed 'decoder.'."""
    sd: Dict[str, torch.Tensor] = {}

    def conv(name, co, ci, k):
        std = 1.0 / math.sqrt(ci * k * k)
        sd[f"decoder.{name}.weight"] = _normal(seed, name + ".w", (co, ci, k, k), std)
        sd[f"decoder.{name}.bias"] = _normal(seed, name + ".b", (co,), 0.02)

    def norm(name, c):
        sd[f"decoder.{name}.weight"] = 1.0 + _normal(seed, name + ".w", (c,), 0.1)
        sd[f"decoder.{name}.bias"] = _normal(seed, name + ".b", (c,), 0.05)

    for kind, key, ci, co in vq_decoder_plan():
        if kind == "conv1":
            conv(key, co, ci, 1)
        elif kind == "conv3":
            conv(key, co, ci, 3)
        elif kind == "res":
            norm(key + ".norm1", ci); conv(key + ".conv1", co, ci, 3)
            norm(key + ".norm2", co); conv(key + ".conv2", co, co, 3)
            if ci != co:
                conv(key + ".nin_shortcut", co, ci, 1)
        elif kind == "attn":
            norm(key + ".norm", ci)
            for n in ("q", "k", "v", "proj_out"):
                conv(f"{key}.{n}", co, ci, 1)
        elif kind == "up":
            conv(key + ".conv", co, ci, 3)
        elif kind == "norm_out":
            norm(key, ci)
    return sd


# --- MAGVIT-v2 encoder --------------------------------------------------------------------

VQ_ENC_CH_MULT, VQ_ENC_NUM_RES = (1, 2, 2, 4, 4), (4, 3, 4, 3, 4)


def vq_encoder_plan():
    """Encoder topology as a flat list of (kind, key-prefix, c_in, c_out), following reference
    models/modeling_magvitv2.py:74-141 (construction) / :143-169 (forward).  attn_resolutions [5] never
    matches a feature-map size, so the down path has no attention blocks."""
    plan = [("conv3", "conv_in", 3, VQ_CH)]
    c = VQ_CH
    for lvl in range(5):
        co = VQ_CH * VQ_ENC_CH_MULT[lvl]
        for j in range(VQ_ENC_NUM_RES[lvl]):
            plan.append(("res", f"down.{lvl}.block.{j}", c, co))
            c = co
        if lvl != 4:
            plan.append(("down", f"down.{lvl}.downsample", c, c))
    plan += [("res", "mid.block_1", c, c), ("attn", "mid.attn_1", c, c), ("res", "mid.block_2", c, c),
             ("norm_out", "norm_out", c, c), ("conv3", "conv_out", c, VQ_Z), ("conv1", "quant_conv", VQ_Z, VQ_Z)]
    return plan


def make_vq_encoder_weights(seed: int = 0) -> Dict[str, torch.Tensor]:
    """fp32 state dict for ``MAGVITv2.encoder`` with the reference's key names, prefixed 'encoder.'."""
    sd: Dict[str, torch.Tensor] = {}

    def conv(name, co, ci, k):
        std = 1.0 / math.sqrt(ci * k * k)
        sd[f"encoder.{name}.weight"] = _normal(seed, "enc." + name + ".w", (co, ci, k, k), std)
        sd[f"encoder.{name}.bias"] = _normal(seed, "enc." + name + ".b", (co,), 0.02)

    def norm(name, c):
        sd[f"encoder.{name}.weight"] = 1.0 + _normal(seed, "enc." + name + ".w", (c,), 0.1)
        sd[f"encoder.{name}.bias"] = _normal(seed, "enc." + name + ".b", (c,), 0.05)

    for kind, key, ci, co in vq_encoder_plan():
        if kind == "conv1":
            conv(key, co, ci, 1)
        elif kind == "conv3":
            conv(key, co, ci, 3)
        elif kind == "res":
            norm(key + ".norm1", ci); conv(key + ".conv1", co, ci, 3)
            norm(key + ".norm2", co); conv(key + ".conv2", co, co, 3)
            if ci != co:
                conv(key + ".nin_shortcut", co, ci, 1)
        elif kind == "attn":
            norm(key + ".norm", ci)
            for n in ("q", "k", "v", "proj_out"):
                conv(f"{key}.{n}", co, ci, 1)
        elif kind == "down":
            conv(key + ".conv", co, ci, 3)
        elif kind == "norm_out":
            norm(key, ci)
    return sd

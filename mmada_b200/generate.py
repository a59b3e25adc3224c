"""LLaDA-style semi-autoregressive text generation with low-confidence remasking — mirror of the
reference's free function ``generate`` (/root/reference/generate.py:43-113) and its helpers
``add_gumbel_noise`` (:8-19) and ``get_num_transfer_tokens`` (:22-40), on the B200 kernels.

Per step the reference materialises fp64 Gumbel noise and an fp64 softmax over the whole (B, L, V)
logits tensor.  Only still-masked positions of the current block can be transferred (every
earlier block is complete, every later position is forced to -inf), so
this implementation computes ln_f + the vocabulary projection + the fp64 sampling chain on the block's
rows only.  Consequence, stated plainly: with ``temperature > 0`` the uniforms are by default drawn per
candidate row (in-kernel Philox, or ``noise=`` for parity tests), not as one (B, L, V) draw from the
global generator, so the random stream differs from the reference's; given the same uniforms the
decisions are identical.  ``rng="reference"`` makes the reference's own draw instead — one
``torch.rand((B, L, V), dtype=float64)`` per forward from the device's global generator (generate.py:14:
``torch.rand_like(logits, dtype=torch.float64)``), of which the block's rows are used — so a seeded generator
reproduces the reference's stream; it costs the reference's memory for that tensor (B·L·V·8 bytes).
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import ops


def get_num_transfer_tokens(mask_index: torch.Tensor, steps: int) -> torch.Tensor:
    """(B, block) bool -> (B, steps) int64: mask_num // steps, +1 for the first mask_num % steps steps."""
    mask_num = mask_index.sum(dim=1, keepdim=True)
    base = mask_num // steps
    remainder = mask_num % steps
    ar = torch.arange(steps, device=mask_index.device)[None, :]
    return (base + (ar < remainder).to(torch.int64)).to(torch.int64)


def _model_device(model):
    dev = getattr(model, "device", None)
    return torch.device("cuda") if dev is None else torch.device(dev)


@torch.no_grad()
def generate(model, prompt, steps=128, gen_length=128, block_length=128, temperature=0.,
             cfg_scale=0., remasking='low_confidence', mask_id=126336, attention_mask=None, *,
             noise: Optional[Sequence[torch.Tensor]] = None, eot_token: Optional[int] = None, seed: Optional[int] = None,
             trace: Optional[list] = None, stop_after_steps: Optional[int] = None, rng: str = "philox"):
    """Same positional/keyword arguments as the reference.  ``model`` must be an
    ``mmada_b200.LLaDAModelLM`` (it provides ``logits_rows``).  ``attention_mask`` is accepted and has
    no effect, as in the reference (the bias built from it is never applied, Q1).  Precondition: ``prompt`` holds no
    ``mask_id`` (raises otherwise; the reference would also unmask masked prompt positions).  Keyword-only
    extras: ``noise`` (per forward k, fp64 uniforms (B, block_length, V) for the block's rows),
    ``eot_token`` (mmu_generate_fast's early exit), ``seed`` for the in-kernel generator, ``stop_after_steps``
    (benchmarking hook: return after that many forwards, sequence length unchanged)."""
    if remasking not in ('low_confidence', 'random'):
        raise NotImplementedError(remasking)
    if rng not in ("philox", "reference"):
        raise ValueError(f"rng must be 'philox' or 'reference', not {rng!r}")
    dev = _model_device(model)
    prompt = prompt.to(dev)
    B, Lp = prompt.shape
    # Precondition of the block restriction (module docstring, Q21): the prompt holds no mask_id.  The reference ranks
    # every still-masked position before the block end (generate.py:98-111), so a masked PROMPT position would be a
    # transfer candidate there and never here — refuse instead of diverging silently (one host read per call).
    if bool((prompt == mask_id).any()):
        raise NotImplementedError("generate(): the prompt contains mask_id tokens (infilling-style input); this "
                                  "implementation only ranks the still-masked positions of the current block")
    L = Lp + gen_length
    x = torch.full((B, L), mask_id, dtype=torch.long, device=dev)
    x[:, :Lp] = prompt.clone()
    prompt_index = (x != mask_id)
    assert gen_length % block_length == 0
    num_blocks = gen_length // block_length
    assert steps % num_blocks == 0
    steps = steps // num_blocks
    if seed is None:
        seed = int(torch.empty((), dtype=torch.int64).random_().item())      # host RNG, no device sync
    cfg = cfg_scale > 0.
    R = 2 * B if cfg else B
    rows_base = (torch.arange(R, device=dev, dtype=torch.int32)[:, None] * L
                 + torch.arange(block_length, device=dev, dtype=torch.int32)[None, :])
    k = 0
    for num_block in range(num_blocks):
        lo = Lp + num_block * block_length
        cnt = ops.block_mask_count(x, lo, block_length, mask_id)
        rows = (rows_base + lo).reshape(-1).contiguous()
        for i in range(steps):
            if stop_after_steps is not None and k >= stop_after_steps:
                return x
            if cfg:
                un_x = x.clone()
                un_x[prompt_index] = mask_id
                x_ = torch.cat([x, un_x], dim=0)
            else:
                x_ = x
            logits = model.logits_rows(x_, rows)                              # [R*block, V] fp32
            n = B * block_length
            u = None
            if noise is not None and temperature != 0:
                u = noise[k].to(dev).reshape(n, -1).contiguous()
            elif rng == "reference" and temperature != 0:
                # the reference's draw: every position of the (B, L, V) logits gets a uniform; the block's rows are read
                V = logits.shape[-1]
                u = torch.rand((B, L, V), dtype=torch.float64, device=dev)[:, lo:lo + block_length].reshape(n, V).contiguous()
            x0, conf = ops.text_sample_rows(logits[:n], logits[n:] if cfg else None, cfg_scale, temperature, u,
                                            seed=seed + k)
            override = None
            if remasking == 'random':
                # the reference's draw (generate.py:90, modeling_mmada.py:455): (B, L) fp32 uniforms from the device's
                # global generator — same shape, dtype and order, so a seeded generator reproduces its stream; only the
                # current block's columns can be selected (the rest is -inf there)
                override = torch.rand((B, L), device=dev)[:, lo:lo + block_length].to(torch.float64).contiguous()
            tr = ops.text_transfer(x, lo, block_length, x0, conf, cnt, steps, i, mask_id, conf_override=override,
                                   want_transfer=trace is not None)
            if hasattr(model, "kernel_launches"):
                model.kernel_launches += 2
            if trace is not None:
                trace.append(dict(k=k, block=num_block, step=i, x0=x0.view(B, block_length).clone(),
                                  conf=conf.view(B, block_length).clone(), transfer=tr, x=x.clone(),
                                  logits=logits.clone(), override=override, u=u))     # (a graph replay reuses the buffer)
            k += 1
        if eot_token is not None:
            last = lo + block_length - 1
            if last < x.shape[1] and bool(torch.all(x[:, last] == eot_token)):
                break
    return x

"""Oracle: the masked-diffusion denoising loops and their sampling math, restated.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
  * /root/reference/models/sampling.py:10-16,31-40      log / gumbel_noise / mask_by_random_topk / cosine_schedule
  * /root/reference/models/modeling_mmada.py:117-211    MMadaModelLM.t2i_generate
  * /root/reference/generate.py:8-40,43-113             add_gumbel_noise / get_num_transfer_tokens / generate
    (== modeling_mmada.py:388-481 mmu_generate, :483-556 mmu_generate_fast)
with the behavioural quirks of SURVEY.md Appendix A (Q1-Q21) kept: no attention bias, compounding
temperature, host fp32 cosine, strict `<` at the cut-off, multinomial == argmax(p / q) with
q ~ Exp(1) drawn before u ~ U(0,1).

Every function takes the noise either from a ``torch.Generator`` exactly as the reference draws it,
or explicitly (``noise=``) so that a CUDA implementation can be fed the same tensors.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F


# ---- models/sampling.py --------------------------------------------------------------------

def cosine_schedule(t: torch.Tensor) -> torch.Tensor:
    return torch.cos(t * math.pi * 0.5)


def _safe_log(t: torch.Tensor, eps: float = 1e-20) -> torch.Tensor:
    return torch.log(t.clamp(min=eps))


def gumbel_from_uniform(u: torch.Tensor) -> torch.Tensor:
    return -_safe_log(-_safe_log(u))


def mask_by_random_topk(mask_len: torch.Tensor, probs: torch.Tensor, temperature: float = 1.0,
                        generator: Optional[torch.Generator] = None,
                        u: Optional[torch.Tensor] = None) -> torch.Tensor:
    if u is None:
        u = torch.zeros_like(probs).uniform_(0, 1, generator=generator)
    confidence = _safe_log(probs) + temperature * gumbel_from_uniform(u)
    sorted_confidence = torch.sort(confidence, dim=-1).values
    cut_off = torch.gather(sorted_confidence, 1, mask_len.long())
    return confidence < cut_off


def t2i_mask_len_schedule(n_tokens: int, timesteps: int, schedule: Callable = cosine_schedule) -> List[float]:
    """floor(N * schedule((s+1)/T)) for s = 0..T-1, computed like modeling_mmada.py:186-195 (a 0-d CPU
    fp32 tensor).  Appendix C: differs from float64 arithmetic at s=9 and s=T-1."""
    out = []
    for s in range(timesteps):
        ratio = 1.0 * (s + 1) / timesteps
        out.append(float((n_tokens * schedule(torch.tensor(ratio))).floor()))
    return out


# ---- MMadaModelLM.t2i_generate -------------------------------------------------------------

def t2i_sample_step(cond_logits: torch.Tensor, uncond_logits: Optional[torch.Tensor], guidance_scale: float,
                    known_ids: torch.Tensor, mask_token_id: int, mask_len_raw: float, temperature: float,
                    q: torch.Tensor, u: torch.Tensor) -> Dict[str, torch.Tensor]:
    """One sampling step on already-sliced logits (B, N, C).  modeling_mmada.py:164-209.
    ``known_ids`` is `input_ids_minus_lm_vocab_size` (B, N): code ids, or mask_token_id where unknown.
    ``q`` (B*N, C) ~ Exp(1), ``u`` (B, N) ~ U(0,1).  ``temperature`` is the already-annealed value."""
    if uncond_logits is not None:
        logits = (1 + guidance_scale) * cond_logits - guidance_scale * uncond_logits
    else:
        logits = cond_logits
    B, N, C = logits.shape
    probs = logits.softmax(dim=-1)
    ratio = probs.reshape(-1, C) / q
    sampled = torch.argmax(ratio, dim=-1).view(B, N)
    unknown = known_ids == mask_token_id
    sampled = torch.where(unknown, sampled, known_ids)
    sel = torch.gather(probs, -1, sampled.long()[..., None]).squeeze(-1)
    sel = torch.where(unknown, sel, torch.finfo(sel.dtype).max)
    mask_len = torch.tensor(mask_len_raw).unsqueeze(0).to(logits.device)
    mask_len = torch.max(torch.tensor([1], device=logits.device),
                         torch.min(unknown.sum(dim=-1, keepdim=True) - 1, mask_len))
    masking = mask_by_random_topk(mask_len, sel, temperature, u=u)
    return dict(sampled_ids=sampled, selected_probs=sel, masking=masking, mask_len=mask_len,
                next_known=torch.where(masking, mask_token_id, sampled))


def t2i_generate(logits_fn: Callable[[torch.Tensor], torch.Tensor],
                 input_ids: torch.Tensor, uncond_input_ids: Optional[torch.Tensor] = None, *,
                 temperature: float = 1.0, timesteps: int = 18, guidance_scale: float = 0,
                 noise_schedule: Callable = cosine_schedule, generator: Optional[torch.Generator] = None,
                 seq_len: int = 1024, mask_token_id: int = 126336, resolution: int = 512,
                 codebook_size: int = 8192, text_vocab: int = 126349,
                 noise: Optional[Sequence[Tuple[torch.Tensor, torch.Tensor]]] = None,
                 sliced_logits: bool = False, trace: Optional[list] = None) -> torch.Tensor:
    """``logits_fn(ids (R, L)) -> (R, L, V)`` full logits, or, when ``sliced_logits``, already the
    (R, N, codebook) block.  Mutates ``input_ids`` in place like the reference (:206)."""
    N = seq_len
    known = input_ids[:, -(N + 1):-1].clone()
    known = torch.where(known == mask_token_id, mask_token_id, known - text_vocab)
    cfg = uncond_input_ids is not None and guidance_scale > 0
    if uncond_input_ids is not None:
        uncond_prefix = uncond_input_ids[:, :resolution + 1]
    sampled = None
    for step in range(timesteps):
        if cfg:
            uncond_input_ids = torch.cat([uncond_prefix, input_ids[:, resolution + 1:]], dim=1)
            logits = logits_fn(torch.cat([input_ids, uncond_input_ids]))
            cond, unc = torch.chunk(logits, 2, dim=0)
            if not sliced_logits:
                # the reference mixes on the full tensor and slices afterwards; elementwise, so equal
                cond = cond[:, -(N + 1):-1, text_vocab:text_vocab + codebook_size]
                unc = unc[:, -(N + 1):-1, text_vocab:text_vocab + codebook_size]
        else:
            cond, unc = logits_fn(input_ids), None
            if not sliced_logits:
                cond = cond[:, -(N + 1):-1, text_vocab:text_vocab + codebook_size]
        B = cond.shape[0]
        if noise is not None:
            q, u = noise[step]
        else:
            q = torch.empty(B * N, codebook_size, dtype=cond.dtype, device=cond.device).exponential_(1, generator=generator)
            u = None
        ratio = 1.0 * (step + 1) / timesteps
        mask_len_raw = float((N * noise_schedule(torch.tensor(ratio))).floor())
        temperature = temperature * (1.0 - ratio)
        if u is None:
            # RNG order (Q11): q first, then u
            u = torch.zeros(B, N, dtype=cond.dtype, device=cond.device).uniform_(0, 1, generator=generator)
        r = t2i_sample_step(cond, unc, guidance_scale, known, mask_token_id, mask_len_raw, temperature, q, u)
        sampled = r["sampled_ids"]
        input_ids[:, -(N + 1):-1] = torch.where(r["masking"], mask_token_id, sampled + text_vocab)
        known = r["next_known"]
        if trace is not None:
            trace.append(dict(step=step, cond=cond, uncond=unc, q=q, u=u, temperature=temperature,
                              mask_len_raw=mask_len_raw, **r))
    return sampled


# ---- MMadaModelLM.t2m_generate (models/modelling_ours.py:557-682) ---------------------------------

def t2m_generate(logits_fn: Callable[[torch.Tensor], torch.Tensor], input_ids: torch.Tensor, *, temperature: float = 1.0,
                 timesteps: int = 18, noise_schedule: Callable = cosine_schedule,
                 generator: Optional[torch.Generator] = None, seq_len: int = 256, mask_token_id: int = 126336,
                 motion_vocab_size: int = 512, text_vocab: int = 126349, image_codebook_size: int = 8192,
                 som_token: Optional[int] = None, eom_token: Optional[int] = None,
                 noise: Optional[Sequence[Tuple[torch.Tensor, Optional[torch.Tensor]]]] = None,
                 trace: Optional[list] = None) -> torch.Tensor:
    """No CFG, non-compounding temperature, no re-masking on the last step, returns the last step's raw
    samples (Q15).  Mutates ``input_ids`` in place (offset token ids)."""
    start = end = None
    if som_token is not None:
        pos = (input_ids == som_token).nonzero(as_tuple=True)
        if len(pos[1]) > 0:
            start = pos[1][0].item() + 1
    if eom_token is not None:
        pos = (input_ids == eom_token).nonzero(as_tuple=True)
        if len(pos[1]) > 0:
            end = pos[1][0].item()
    if start is None or end is None:
        start, end = input_ids.shape[1] - seq_len, input_ids.shape[1]
    local = input_ids[:, start:end].clone()
    off = text_vocab + image_codebook_size
    sampled_ids = None
    for step in range(timesteps):
        logits = logits_fn(input_ids)
        ml = logits[:, start:end, off:off + motion_vocab_size]
        probs = ml.softmax(dim=-1)
        flat = probs.reshape(-1, ml.size(-1))
        if noise is not None:
            q, u = noise[step]
        else:
            q, u = torch.empty_like(flat).exponential_(1, generator=generator), None
        sampled_ids = torch.argmax(flat / q, dim=-1).view(*ml.shape[:-1])
        merged = sampled_ids + off
        unknown = local == mask_token_id
        merged = torch.where(unknown, merged, local)
        input_ids[:, start:end] = merged
        masking = None
        if step < timesteps - 1:
            ratio = 1.0 * (step + 1) / timesteps
            mask_ratio = noise_schedule(torch.tensor(ratio))
            sel = torch.gather(probs, -1, sampled_ids.long()[..., None]).squeeze(-1)
            sel = torch.where(unknown, sel, torch.finfo(sel.dtype).max)
            mask_len = (seq_len * mask_ratio).floor().unsqueeze(0).to(ml.device)
            mask_len = torch.max(torch.tensor([1], device=ml.device), torch.min(unknown.sum(dim=-1, keepdim=True) - 1, mask_len))
            masking = mask_by_random_topk(mask_len, sel, temperature * (1.0 - ratio), generator=generator, u=u)
            local = torch.where(masking, mask_token_id, merged)
            input_ids[:, start:end] = torch.where(masking, mask_token_id, merged)
        if trace is not None:
            trace.append(dict(step=step, logits=ml, q=q, u=u, sampled_ids=sampled_ids, merged=merged, masking=masking))
    return sampled_ids


# ---- generate.py / mmu_generate -------------------------------------------------------------

def add_gumbel_noise(logits: torch.Tensor, temperature: float, u: Optional[torch.Tensor] = None) -> torch.Tensor:
    if temperature == 0:
        return logits
    logits = logits.to(torch.float64)
    if u is None:
        u = torch.rand_like(logits, dtype=torch.float64)
    return logits.exp() / ((-torch.log(u)) ** temperature)


def get_num_transfer_tokens(mask_index: torch.Tensor, steps: int) -> torch.Tensor:
    mask_num = mask_index.sum(dim=1, keepdim=True)
    base, rem = mask_num // steps, mask_num % steps
    out = torch.zeros(mask_num.size(0), steps, device=mask_index.device, dtype=torch.int64) + base
    for i in range(mask_num.size(0)):
        out[i, :rem[i]] += 1
    return out


def text_sample_rows(logits: torch.Tensor, temperature: float, u: Optional[torch.Tensor]):
    """Per-row Gumbel-max token and its fp64 softmax probability.  generate.py:90-96.
    logits (R, V) model dtype; u (R, V) fp64 or None when temperature == 0."""
    x0 = torch.argmax(add_gumbel_noise(logits, temperature, u), dim=-1)
    p = F.softmax(logits.to(torch.float64), dim=-1)
    return x0, torch.gather(p, -1, x0[:, None]).squeeze(-1)


def generate(logits_fn: Callable[[torch.Tensor], torch.Tensor], prompt: torch.Tensor, steps: int = 128,
             gen_length: int = 128, block_length: int = 128, temperature: float = 0.0, cfg_scale: float = 0.0,
             remasking: str = "low_confidence", mask_id: int = 126336,
             noise: Optional[Sequence[torch.Tensor]] = None, eot_token: Optional[int] = None,
             trace: Optional[list] = None) -> torch.Tensor:
    """``logits_fn(x (R, L)) -> (R, L, V)``.  noise[k] (B, L, V) fp64 uniform for forward k, else the
    global RNG like the reference.  ``eot_token`` enables mmu_generate_fast's early exit (:550-555)."""
    B, Lp = prompt.shape
    x = torch.full((B, Lp + gen_length), mask_id, dtype=torch.long, device=prompt.device)
    x[:, :Lp] = prompt.clone()
    prompt_index = x != mask_id
    assert gen_length % block_length == 0
    num_blocks = gen_length // block_length
    assert steps % num_blocks == 0
    steps = steps // num_blocks
    k = 0
    for nb in range(num_blocks):
        lo, hi = Lp + nb * block_length, Lp + (nb + 1) * block_length
        ntt = get_num_transfer_tokens(x[:, lo:hi] == mask_id, steps)
        for i in range(steps):
            mask_index = x == mask_id
            if cfg_scale > 0.0:
                un_x = x.clone()
                un_x[prompt_index] = mask_id
                logits = logits_fn(torch.cat([x, un_x], dim=0))
                logits, un_logits = torch.chunk(logits, 2, dim=0)
                logits = un_logits + (cfg_scale + 1) * (logits - un_logits)
            else:
                logits = logits_fn(x)
            u = None if noise is None or temperature == 0 else noise[k]
            x0 = torch.argmax(add_gumbel_noise(logits, temperature, u), dim=-1)
            if remasking == "low_confidence":
                p = F.softmax(logits.to(torch.float64), dim=-1)
                x0_p = torch.squeeze(torch.gather(p, dim=-1, index=torch.unsqueeze(x0, -1)), -1)
            elif remasking == "random":
                x0_p = torch.rand((x0.shape[0], x0.shape[1]), device=x0.device)
            else:
                raise NotImplementedError(remasking)
            x0_p[:, hi:] = -np.inf
            x0 = torch.where(mask_index, x0, x)
            confidence = torch.where(mask_index, x0_p, -np.inf)
            transfer = torch.zeros_like(x0, dtype=torch.bool)
            for j in range(B):
                _, sel = torch.topk(confidence[j], k=int(ntt[j, i]))
                transfer[j, sel] = True
            x[transfer] = x0[transfer]
            if trace is not None:
                trace.append(dict(k=k, block=nb, step=i, x0=x0.clone(), confidence=confidence.clone(),
                                  transfer=transfer.clone(), x=x.clone()))
            k += 1
        if eot_token is not None:
            last = x[:, hi - 1]
            if (last == eot_token).all():
                break
    return x

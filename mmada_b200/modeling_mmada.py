"""MMadaModelLM — host-side mirror of the reference's denoising loops
(/root/reference/models/modeling_mmada.py:106-211 ``t2i_generate``, :388-481 ``mmu_generate``,
:483-556 ``mmu_generate_fast``, :558-663 ``t2i_generate_decoding_stepwise``) on the B200 kernels.

Same method names, keyword names, defaults, in-place mutation of ``input_ids`` and return values as
the reference; what differs is what runs underneath:
  * one transformer forward per step on [cond ; uncond] rows through the tcgen05 GEMM / attention
    kernels (modeling_llada.LLaDAModelLM), never building the attention bias the reference builds and
    ignores (Q1/Q2);
  * ln_f + output head only on the image positions and the codebook columns (Q5/Q16), fp32 logits;
  * one fused sampling kernel per step (csrc/sampling.cu) instead of ~25 eager ops; no host syncs
    inside the loop (the reference has a .item() at entry and a CPU->GPU copy per step, Q10).
Noise is drawn with the same torch calls, shapes and order as the reference (Q6/Q11), so a given
``generator`` yields the reference's own noise stream.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from . import generate as _generate
from .modeling_llada import CausalLMOutput, LLaDAConfig, LLaDAModelLM
from .sampling import cosine_schedule


class MMadaConfig(LLaDAConfig):
    """Config of the reference's MMadaConfig/ModelConfig that the path reads."""


class MMadaModelLM(LLaDAModelLM):
    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def t2i_generate(
            self,
            input_ids: torch.LongTensor = None,
            uncond_input_ids: torch.LongTensor = None,
            attention_mask=None,
            uncond_attention_mask=None,
            temperature=1.0,
            timesteps=18,
            guidance_scale=0,
            noise_schedule=cosine_schedule,
            generator: torch.Generator = None,
            config=None,
            seq_len=1024,
            mask_token_id=126336,
            resolution=512,
            codebook_size=8192,
            **kwargs,
    ):
        """MaskGIT-style parallel decoding with classifier-free guidance; returns (B, seq_len) int64
        code ids and leaves ``input_ids`` mutated like the reference (:206).  Extra kwargs:
        ``uni_prompting`` (only ``len(uni_prompting.text_tokenizer)`` is read, :149) and, for parity
        tests, ``noise`` = per-step list of (q [B*N, C] ~ Exp(1), u [B, N] ~ U(0,1))."""
        uni_prompting = kwargs.get("uni_prompting", None)
        text_vocab = len(uni_prompting.text_tokenizer)
        noise = kwargs.get("noise", None)
        trace = kwargs.get("trace", None)
        stop_after = kwargs.get("stop_after_steps", None)     # benchmarking hook: cut the loop short
        N, C = seq_len, codebook_size
        dev = self.device
        caller_ids = input_ids
        input_ids = input_ids if input_ids.is_cuda else input_ids.to(dev)       # the caller's tensor when on device
        B, L = input_ids.shape
        img_off = L - (N + 1)
        known = input_ids[:, img_off:img_off + N].clone()
        known = torch.where(known == mask_token_id, mask_token_id, known - text_vocab).contiguous()
        cfg = uncond_input_ids is not None and guidance_scale > 0
        R = 2 * B if cfg else B
        model_input = torch.empty((R, L), dtype=torch.int64, device=dev)
        if cfg:
            P = resolution + 1                                                  # text-prefix length (Q4)
            model_input[B:, :P] = uncond_input_ids.to(dev)[:, :P]
        # flattened token rows of the image positions, cond rows first then uncond rows
        rows = (torch.arange(R, device=dev, dtype=torch.int32)[:, None] * L + img_off
                + torch.arange(N, device=dev, dtype=torch.int32)[None, :]).reshape(-1).contiguous()
        tickets = torch.zeros(B, dtype=torch.int32, device=dev)
        sampled = None
        for step in range(timesteps):
            if stop_after is not None and step >= stop_after:
                break
            model_input[:B] = input_ids
            if cfg:
                model_input[B:, P:] = input_ids[:, P:]
            logits = self.logits_rows(model_input, rows, text_vocab, text_vocab + C)      # [R*N, C] fp32
            cond = logits[:B * N]
            unc = logits[B * N:] if cfg else None
            if noise is not None:
                q, u = noise[step]
                q, u = q.to(dev), u.to(dev)
            else:
                q = torch.empty((B * N, C), dtype=torch.float32, device=dev).exponential_(1, generator=generator)
                u = None
            ratio = 1.0 * (step + 1) / timesteps
            mask_ratio = noise_schedule(torch.tensor(ratio))                    # host fp32, like :187
            mask_len_raw = float((N * mask_ratio).floor())
            temperature = temperature * (1.0 - ratio)                           # compounding (Q3)
            if u is None:
                u = torch.zeros((B, N), dtype=torch.float32, device=dev).uniform_(0, 1, generator=generator)
            sampled, sel, masking = ops.t2i_sample_step(cond, unc, q, u, known, input_ids, img_off, tickets,
                                                        guidance_scale if cfg else 0.0, mask_len_raw, temperature,
                                                        mask_token_id, text_vocab, want_masking=trace is not None)
            self.kernel_launches += 1
            if trace is not None:
                trace.append(dict(step=step, cond=cond.view(B, N, C).clone(), uncond=None if unc is None else unc.view(B, N, C).clone(),
                                  sampled_ids=sampled, selected_probs=sel, masking=masking))
        if caller_ids is not input_ids:
            caller_ids.copy_(input_ids)                                          # keep the in-place contract
        return sampled

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def mmu_generate(self, idx=None, input_embeddings=None, max_new_tokens=128, steps=128, block_length=128,
                     temperature=0.0, top_k=None, eot_token=None, cfg_scale=0.0, remasking='low_confidence',
                     mask_id=126336, attention_mask=None, **kwargs):
        """Reference models/modeling_mmada.py:388-481: the same algorithm as ``generate()``.  ``top_k``,
        ``input_embeddings`` and ``eot_token`` are accepted and ignored like the reference (Q14)."""
        return _generate.generate(self, idx, steps=steps, gen_length=max_new_tokens, block_length=block_length,
                                  temperature=temperature, cfg_scale=cfg_scale, remasking=remasking, mask_id=mask_id,
                                  attention_mask=attention_mask, **kwargs)

    @torch.no_grad()
    def mmu_generate_fast(self, idx=None, input_embeddings=None, max_new_tokens=128, steps=128, block_length=128,
                          temperature=0.0, top_k=None, eot_token=None, cfg_scale=0.0, remasking='low_confidence',
                          mask_id=126336, attention_mask=None, **kwargs):
        """Reference :483-556: ``mmu_generate`` plus an early exit once every row's block-final token is
        ``eot_token`` (one device->host read per block, as in the reference)."""
        return _generate.generate(self, idx, steps=steps, gen_length=max_new_tokens, block_length=block_length,
                                  temperature=temperature, cfg_scale=cfg_scale, remasking=remasking, mask_id=mask_id,
                                  attention_mask=attention_mask, eot_token=eot_token, **kwargs)

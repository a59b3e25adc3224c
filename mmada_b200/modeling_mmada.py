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
from .generate import generate as _generate_fn
from .modeling_llada import CausalLMOutput, LLaDAConfig, LLaDAModelLM
from .sampling import cosine_schedule


class MMadaConfig(LLaDAConfig):
    """Config of the reference's MMadaConfig/ModelConfig that the path reads."""


class MMadaModelLM(LLaDAModelLM):
    @staticmethod
    def _config_class():
        return MMadaConfig

    # ------------------------------------------------------------------------------------------
    def _t2i_steps(self, input_ids, uncond_input_ids, temperature, timesteps, guidance_scale, noise_schedule, generator,
                   seq_len, mask_token_id, resolution, codebook_size, kwargs):
        """The denoising loop of t2i_generate as a Python generator: yields (step, sampled_ids) after
        every step (sampled_ids = the step's predictions merged with the already-known tokens)."""
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
        # The last block, ln_f and the output head run on the STILL-MASKED positions only (the reference computes and
        # discards the logits of known positions, :183-184).  `cap` bounds the masked positions per row without a host
        # read: N at the first step (the caller may pass known tokens), then the previous step's mask_len.
        cap = N
        for step in range(timesteps):
            if stop_after is not None and step >= stop_after:
                break
            model_input[:B] = input_ids
            if cfg:
                model_input[B:, P:] = input_ids[:, P:]
            slot = None
            if self.masked_rows_only and cap < N:
                step_rows, slot = ops.compact_masked_rows(known, L, img_off, cap, 2 if cfg else 1, mask_token_id)
                self.kernel_launches += 1
            else:
                step_rows = rows
            logits = self.logits_rows(model_input, step_rows, text_vocab, text_vocab + C)     # [R*cap, C] fp32
            half = logits.shape[0] // (2 if cfg else 1)
            cond = logits[:half]
            unc = logits[half:] if cfg else None
            if noise is not None:
                q, u = noise[step]
                q, u = q.to(dev), u.to(dev)
            elif isinstance(generator, (list, tuple)):
                # one generator per prompt: the noise of a prompt does not depend on batch composition / world size
                q = torch.empty((B * N, C), dtype=torch.float32, device=dev)
                u = torch.empty((B, N), dtype=torch.float32, device=dev)
                for bi, gb in enumerate(generator):
                    q[bi * N:(bi + 1) * N].exponential_(1, generator=gb)
                    u[bi].uniform_(0, 1, generator=gb)
            else:
                q = torch.empty((B * N, C), dtype=torch.float32, device=dev).exponential_(1, generator=generator)
                u = None
            ratio = 1.0 * (step + 1) / timesteps
            mask_ratio = noise_schedule(torch.tensor(ratio))                    # host fp32, like :187
            mask_len_raw = float((N * mask_ratio).floor())
            temperature = temperature * (1.0 - ratio)                           # compounding (Q3)
            if u is None:
                u = torch.zeros((B, N), dtype=torch.float32, device=dev).uniform_(0, 1, generator=generator)
            if trace is not None and slot is not None:
                # parity tests replay the step on the CPU oracle with logits at EVERY position: known positions (whose
                # samples the reference discards) get zeros
                pos = (slot.view(-1) >= 0).nonzero(as_tuple=True)[0]
                src = slot.view(-1)[pos].long()
                full_c = torch.zeros((B * N, C), dtype=torch.float32, device=dev)
                full_c[pos] = cond[src]
                full_u = None
                if unc is not None:
                    full_u = torch.zeros((B * N, C), dtype=torch.float32, device=dev)
                    full_u[pos] = unc[src]
            sampled, sel, masking = ops.t2i_sample_step(cond, unc, q, u, known, input_ids, img_off, tickets,
                                                        guidance_scale if cfg else 0.0, mask_len_raw, temperature,
                                                        mask_token_id, text_vocab, want_masking=trace is not None,
                                                        slot=slot)
            self.kernel_launches += 1
            # at most k = max(1, min(unknown - 1, mask_len)) positions stay masked (sampling.cu, modeling_mmada.py:195-200)
            cap = max(1, min(cap - 1, int(mask_len_raw)))
            if trace is not None:
                if slot is None:
                    full_c, full_u = cond, unc
                trace.append(dict(step=step, cond=full_c.view(B, N, C).clone(),
                                  uncond=None if full_u is None else full_u.view(B, N, C).clone(),
                                  sampled_ids=sampled, selected_probs=sel, masking=masking))
            if caller_ids is not input_ids:
                caller_ids.copy_(input_ids)                                      # keep the in-place contract
            yield step, sampled

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
        tests, ``noise`` = per-step list of (q [B*N, C] ~ Exp(1), u [B, N] ~ U(0,1)).  The attention
        masks are accepted and have no effect, as in the reference (Q1)."""
        sampled = None
        for _, sampled in self._t2i_steps(input_ids, uncond_input_ids, temperature, timesteps, guidance_scale,
                                          noise_schedule, generator, seq_len, mask_token_id, resolution, codebook_size,
                                          kwargs):
            pass
        return sampled

    def t2i_generate_decoding_stepwise(
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
            vq_model=None,
            **kwargs,
    ):
        """Reference models/modeling_mmada.py:558-663: ``t2i_generate`` as a Python generator that also
        decodes the current prediction of row 0 every step and yields ``(PIL.Image, "Step i/T")``.
        (The reference decodes the whole batch and keeps image 0, :627-635; only row 0 is decoded here.)"""
        from PIL import Image
        with torch.no_grad():
            for step, sampled in self._t2i_steps(input_ids, uncond_input_ids, temperature, timesteps, guidance_scale,
                                                 noise_schedule, generator, seq_len, mask_token_id, resolution,
                                                 codebook_size, kwargs):
                codes = torch.clamp(sampled[:1], 0, 8192 - 1)
                if hasattr(vq_model, "decode_code_uint8"):
                    img = vq_model.decode_code_uint8(codes)[0].cpu().numpy()
                else:                                    # any object with the reference's decode_code
                    x = torch.clamp((vq_model.decode_code(codes) + 1.0) / 2.0, min=0.0, max=1.0) * 255.0
                    img = x.permute(0, 2, 3, 1).cpu().numpy().astype("uint8")[0]
                yield Image.fromarray(img), f"Step {step + 1}/{timesteps}"

    @torch.no_grad()
    def t2m_generate(
            self,
            input_ids: torch.LongTensor = None,
            attention_mask=None,
            temperature=1.0,
            timesteps=18,
            noise_schedule=cosine_schedule,
            generator: torch.Generator = None,
            config=None,
            seq_len=256,
            mask_token_id=126336,
            motion_vocab_size=512,
            num_new_special_tokens=0,
            **kwargs,
    ):
        """Text-to-motion variant (reference models/modelling_ours.py:557-682): no CFG, motion vocabulary
        slice [len(tokenizer)+image_codebook_size, +motion_vocab_size), non-compounding temperature, no
        re-masking on the last step, returns the LAST step's raw samples (not merged with kept tokens,
        Appendix A Q15).  ``input_ids`` is updated in place with offset token ids."""
        uni_prompting = kwargs.get("uni_prompting", None)
        noise = kwargs.get("noise", None)
        dev = self.device
        caller_ids = input_ids
        input_ids = input_ids if input_ids.is_cuda else input_ids.to(dev)
        B, L = input_ids.shape
        start = end = None
        spt = getattr(uni_prompting, "sptids_dict", None) if uni_prompting is not None else None
        if spt is not None and "<|som|>" in spt and "<|eom|>" in spt:            # one host read, like the reference
            som = (input_ids == int(spt["<|som|>"])).nonzero(as_tuple=True)[1]
            eom = (input_ids == int(spt["<|eom|>"])).nonzero(as_tuple=True)[1]
            if som.numel() > 0:
                start = int(som[0]) + 1
            if eom.numel() > 0:
                end = int(eom[0])
        if start is None or end is None:
            start, end = L - seq_len, L
        N = end - start
        text_vocab = len(uni_prompting.text_tokenizer) if uni_prompting else 126000
        offset = text_vocab + kwargs.get("image_codebook_size", 8192)
        C = motion_vocab_size
        local = input_ids[:, start:end]
        known = torch.where(local == mask_token_id, mask_token_id, local - offset).contiguous()
        rows = (torch.arange(B, device=dev, dtype=torch.int32)[:, None] * L + start
                + torch.arange(N, device=dev, dtype=torch.int32)[None, :]).reshape(-1).contiguous()
        tickets = torch.zeros(B, dtype=torch.int32, device=dev)
        raw = None
        for step in range(timesteps):
            logits = self.logits_rows(input_ids, rows, offset, offset + C)                  # [B*N, C]
            if noise is not None:
                q, u = (t.to(dev) for t in noise[step])
            else:
                q = torch.empty((B * N, C), dtype=torch.float32, device=dev).exponential_(1, generator=generator)
                u = None
            last = step == timesteps - 1
            ratio = 1.0 * (step + 1) / timesteps
            mask_len_raw = float((seq_len * noise_schedule(torch.tensor(ratio))).floor())
            if u is None:
                u = torch.zeros((B, N), dtype=torch.float32, device=dev)
                if not last:                                                               # the last step draws no u
                    u.uniform_(0, 1, generator=generator)
            _, _, _, raw = ops.t2i_sample_step(logits, None, q, u, known, input_ids, start, tickets, 0.0, mask_len_raw,
                                               temperature * (1.0 - ratio), mask_token_id, offset, want_raw=True,
                                               no_remask=last)
            self.kernel_launches += 1
        if caller_ids is not input_ids:
            caller_ids.copy_(input_ids)
        return raw

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def mmu_generate(self, idx=None, input_embeddings=None, max_new_tokens=128, steps=128, block_length=128,
                     temperature=0.0, top_k=None, eot_token=None, cfg_scale=0.0, remasking='low_confidence',
                     mask_id=126336, attention_mask=None, **kwargs):
        """Reference models/modeling_mmada.py:388-481: the same algorithm as ``generate()``.  ``top_k``,
        ``input_embeddings`` and ``eot_token`` are accepted and ignored like the reference (Q14)."""
        return _generate_fn(self, idx, steps=steps, gen_length=max_new_tokens, block_length=block_length,
                                  temperature=temperature, cfg_scale=cfg_scale, remasking=remasking, mask_id=mask_id,
                                  attention_mask=attention_mask, **kwargs)

    @torch.no_grad()
    def mmu_generate_fast(self, idx=None, input_embeddings=None, max_new_tokens=128, steps=128, block_length=128,
                          temperature=0.0, top_k=None, eot_token=None, cfg_scale=0.0, remasking='low_confidence',
                          mask_id=126336, attention_mask=None, **kwargs):
        """Reference :483-556: ``mmu_generate`` plus an early exit once every row's block-final token is
        ``eot_token`` (one device->host read per block, as in the reference)."""
        return _generate_fn(self, idx, steps=steps, gen_length=max_new_tokens, block_length=block_length,
                                  temperature=temperature, cfg_scale=cfg_scale, remasking=remasking, mask_id=mask_id,
                                  attention_mask=attention_mask, eot_token=eot_token, **kwargs)

    # ------------------------------------------------------------------------------------------
    def _loss_rows(self, ids, lab, batch_size_t2i, max_seq_length, segments):
        """Per-row cross-entropy terms of the training-time forwards: one transformer forward, ``ln_f`` + head + one pass of
        ``mmada_cross_entropy_rows_f32`` on the token rows a loss reads only — the t2i positions behind the text prefix
        whose label is not -100 (rows [0, batch_size_t2i)) and the masked positions of every ``segments`` entry (a Python
        slice of batch rows, in the reference's boolean-mask order).  Returns (nll_t2i, [(nll, mask) per segment])."""
        dev = self.device
        B, L = ids.shape
        masked = ids == self.config.mask_token_id                                             # :246
        pos = torch.arange(B * L, device=dev, dtype=torch.int64).view(B, L)
        if batch_size_t2i > 0:
            t2i_sel = torch.zeros_like(masked)
            t2i_sel[:batch_size_t2i, max_seq_length + 1:] = lab[:batch_size_t2i, max_seq_length + 1:] != -100
            rows = [pos[t2i_sel]]
        else:
            rows = [pos[:0, 0]]
        masks = [masked[sl] for sl in segments]
        rows += [pos[sl][m] for sl, m in zip(segments, masks)]
        counts = [r.numel() for r in rows]
        allrows = torch.cat(rows)
        nll = torch.zeros((0,), dtype=torch.float32, device=dev)
        if allrows.numel() > 0:
            lg = self._logits_rows(ids, allrows.to(torch.int32).contiguous())                 # [n_rows, V] fp32 (no graph: the row count changes per batch)
            nll = ops.cross_entropy_rows(lg, lab.view(-1)[allrows], -100)
            self.kernel_launches += 1
            del lg
        parts = torch.split(nll, counts)
        return parts[0], list(zip(parts[1:], masks))

    @staticmethod
    def _t2i_mean(nll_t2i, batch_size_t2i, dev):
        if batch_size_t2i == 0:
            return torch.tensor(0.0, device=dev)                                              # :237
        n = nll_t2i.numel()
        return nll_t2i.sum() / n if n > 0 else torch.tensor(float("nan"), device=dev)         # mean over the targets

    @torch.no_grad()
    def forward_process(self, input_ids, labels, batch_size_t2i=0, batch_size_lm=0, batch_size_mmu=0, max_seq_length=128,
                        p_mask_lm=None, p_mask_mmu=None, answer_lengths=None, t2i_masks=None, answer_lengths_lm=None,
                        return_logits: bool = False):
        """Forward values of the reference's training step (modeling_mmada.py:213-276): ``(logits, loss_t2i, loss_lm,
        loss_mmu)`` for a mixed t2i / lm / mmu batch — same arguments, same reductions, same quirks (oracle/training.py
        lists them), no autograd: this is the evaluation of the losses (validation / logging), the backward pass is out
        of scope.

        What runs underneath: ``_loss_rows``.  The reference's ``attention_bias`` from ``t2i_masks`` is never applied by
        its attention (Q1) and is not built.  ``logits`` is ``None`` unless ``return_logits`` (the reference returns the
        full (B, L, V) tensor; its training scripts discard it)."""
        dev = self.device
        ids, lab = input_ids.to(dev), labels.to(dev)
        B, L = ids.shape
        lo_lm, hi_lm = batch_size_t2i, batch_size_t2i + batch_size_lm
        lo_mmu = (B - batch_size_mmu) if batch_size_mmu > 0 else 0                            # [-0:] is the whole batch (:249)
        nll_t2i, ((nll_lm, masked_lm), (nll_mmu, masked_mmu)) = self._loss_rows(
            ids, lab, batch_size_t2i, max_seq_length, [slice(lo_lm, hi_lm), slice(lo_mmu, None)])
        loss_t2i = self._t2i_mean(nll_t2i, batch_size_t2i, dev)
        n_rows_lm = float(max(min(hi_lm, B) - lo_lm, 0))                                      # logits[lm].shape[0]
        ce_lm = nll_lm / p_mask_lm.to(dev)[masked_lm]                                         # :253-256
        loss_lm = ce_lm.sum() / torch.tensor(n_rows_lm * L, device=dev)                       # :258 (a scalar ...)
        loss_lm = torch.sum(loss_lm / answer_lengths_lm.to(dev)[masked_lm]) / torch.tensor(n_rows_lm, device=dev)   # :262
        ce_mmu = nll_mmu / p_mask_mmu.to(dev)[masked_mmu]                                     # :264-267
        loss_mmu = torch.sum(ce_mmu / answer_lengths.to(dev)[masked_mmu]) / float(B - lo_mmu)  # :268
        logits = self.forward(ids).logits if return_logits else None
        return logits, loss_t2i, loss_lm, loss_mmu

    @torch.no_grad()
    def forward_process_with_r2i(self, input_ids, labels, t2i_masks=None, max_seq_length=128, batch_size_t2i=0,
                                 batch_size_lm=0, batch_size_mmu=0, batch_size_r2i=0, p_mask_lm=None, p_mask_mmu=None,
                                 p_mask_r2i=None, answer_lengths=None, answer_lengths_lm=None, answer_lengths_r2i=None,
                                 return_logits: bool = False):
        """Forward values of modeling_mmada.py:278-356: like ``forward_process`` with a fourth group of rows (r2i) and
        explicit [start, end) row ranges for every group (no ``[-n:]`` slices); returns ``(logits, loss_t2i, loss_lm,
        loss_mmu, loss_r2i)``."""
        dev = self.device
        ids, lab = input_ids.to(dev), labels.to(dev)
        B, L = ids.shape
        s_lm = batch_size_t2i
        e_lm = s_lm + batch_size_lm
        e_mmu = e_lm + batch_size_mmu
        e_r2i = e_mmu + batch_size_r2i
        nll_t2i, ((nll_lm, m_lm), (nll_mmu, m_mmu), (nll_r2i, m_r2i)) = self._loss_rows(
            ids, lab, batch_size_t2i, max_seq_length, [slice(s_lm, e_lm), slice(e_lm, e_mmu), slice(e_mmu, e_r2i)])
        rows_in = lambda a, b: float(max(min(b, B) - min(a, B), 0))                           # logits[a:b].shape[0]
        loss_t2i = self._t2i_mean(nll_t2i, batch_size_t2i, dev)
        n_lm = rows_in(s_lm, e_lm)
        ce_lm = nll_lm / p_mask_lm.to(dev)[m_lm]
        loss_lm = ce_lm.sum() / torch.tensor(n_lm * L, device=dev)                            # :335 (a scalar, then :336)
        loss_lm = torch.sum(loss_lm / answer_lengths_lm.to(dev)[m_lm]) / torch.tensor(n_lm, device=dev)
        ce_mmu = nll_mmu / p_mask_mmu.to(dev)[m_mmu]
        loss_mmu = torch.sum(ce_mmu / answer_lengths.to(dev)[m_mmu]) / torch.tensor(rows_in(e_lm, e_mmu), device=dev)
        ce_r2i = nll_r2i / p_mask_r2i.to(dev)[m_r2i]
        loss_r2i = torch.sum(ce_r2i / answer_lengths_r2i.to(dev)[m_r2i]) / torch.tensor(rows_in(e_mmu, e_r2i), device=dev)
        logits = self.forward(ids).logits if return_logits else None
        return logits, loss_t2i, loss_lm, loss_mmu, loss_r2i

    @torch.no_grad()
    def forward_t2i(self, input_ids, labels, batch_size_t2i=0, max_seq_length=128, t2i_masks=None):
        """Forward value of modeling_mmada.py:359-385: the t2i cross-entropy alone (mean over the labelled image positions
        of the first ``batch_size_t2i`` rows)."""
        dev = self.device
        ids, lab = input_ids.to(dev), labels.to(dev)
        nll_t2i, _ = self._loss_rows(ids, lab, batch_size_t2i, max_seq_length, [])
        n = nll_t2i.numel()
        return nll_t2i.sum() / n if n > 0 else torch.tensor(float("nan"), device=dev)         # (no batch_size_t2i == 0 branch there)

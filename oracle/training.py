"""Oracle: the training-time forward ``MMadaModelLM.forward_process``, restated (forward values only).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
  * /root/reference/models/modeling_mmada.py:213-276   forward_process: logits + t2i / lm / mmu masked cross-entropy
    (SURVEY.md section 8 row f4), :278-356 forward_process_with_r2i, :359-385 forward_t2i.
Quirks kept literally:
  * the ``attention_bias`` built from ``t2i_masks`` (:228-230) is handed to the model and never applied (Appendix A, Q1),
    so it is not built here;
  * ``loss_lm`` is first reduced to a scalar ``sum / (B_lm * L)`` (:258) and THEN divided element-wise by
    ``answer_lengths_lm[masked]`` and summed (:262): the per-token weights are applied to the scalar;
  * ``masked_indices[-batch_size_mmu:]`` / ``logits[-batch_size_mmu:]`` (:249,264): with ``batch_size_mmu == 0`` the slice
    is the WHOLE batch;
  * the lm / mmu terms index with boolean masks (row-major order of the masked positions).
"""
from __future__ import annotations

from typing import Callable, Tuple

import torch
import torch.nn.functional as F


def forward_process(logits_fn: Callable[[torch.Tensor], torch.Tensor], input_ids: torch.Tensor, labels: torch.Tensor,
                    batch_size_t2i: int = 0, batch_size_lm: int = 0, batch_size_mmu: int = 0, max_seq_length: int = 128,
                    p_mask_lm: torch.Tensor = None, p_mask_mmu: torch.Tensor = None, answer_lengths: torch.Tensor = None,
                    t2i_masks: torch.Tensor = None, answer_lengths_lm: torch.Tensor = None,
                    mask_token_id: int = 126336) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """``logits_fn(ids (B, L)) -> (B, L, V)``.  Returns (logits, loss_t2i, loss_lm, loss_mmu) like the reference."""
    logits = logits_fn(input_ids)                                                              # :231
    V = logits.shape[-1]
    if batch_size_t2i == 0:
        loss_t2i = torch.tensor(0.0, device=input_ids.device)                                  # :237
    else:
        loss_t2i = F.cross_entropy(logits[:batch_size_t2i, max_seq_length + 1:].contiguous().view(-1, V),
                                   labels[:batch_size_t2i, max_seq_length + 1:].contiguous().view(-1),
                                   ignore_index=-100)                                          # :240-243
    masked = input_ids == mask_token_id                                                        # :246
    lm = slice(batch_size_t2i, batch_size_t2i + batch_size_lm)
    masked_lm = masked[lm]
    masked_mmu = masked[-batch_size_mmu:]                                                      # :249 (0 -> everything)
    loss_lm = F.cross_entropy(logits[lm][masked_lm].contiguous().view(-1, V),
                              labels[lm][masked_lm].contiguous().view(-1), ignore_index=-100,
                              reduction="none") / p_mask_lm[masked_lm]                         # :253-256
    loss_lm = loss_lm.sum() / (logits[lm].shape[0] * logits[lm].shape[1])                      # :258
    loss_lm = torch.sum(loss_lm / answer_lengths_lm[masked_lm]) / logits[lm].shape[0]          # :262
    loss_mmu = F.cross_entropy(logits[-batch_size_mmu:][masked_mmu].contiguous().view(-1, V),
                               labels[-batch_size_mmu:][masked_mmu].contiguous().view(-1), ignore_index=-100,
                               reduction="none") / p_mask_mmu[masked_mmu]                      # :264-267
    loss_mmu = torch.sum(loss_mmu / answer_lengths[masked_mmu]) / logits[-batch_size_mmu:].shape[0]   # :268
    return logits, loss_t2i, loss_lm, loss_mmu


def forward_process_with_r2i(logits_fn, input_ids, labels, t2i_masks=None, max_seq_length=128, batch_size_t2i=0,
                             batch_size_lm=0, batch_size_mmu=0, batch_size_r2i=0, p_mask_lm=None, p_mask_mmu=None,
                             p_mask_r2i=None, answer_lengths=None, answer_lengths_lm=None, answer_lengths_r2i=None,
                             mask_token_id: int = 126336):
    """/root/reference/models/modeling_mmada.py:278-356: forward_process with a fourth group (r2i) and explicit row ranges."""
    logits = logits_fn(input_ids)
    V = logits.shape[-1]
    if batch_size_t2i == 0:
        loss_t2i = torch.tensor(0.0, device=input_ids.device)
    else:
        loss_t2i = F.cross_entropy(logits[:batch_size_t2i, max_seq_length + 1:].contiguous().view(-1, V),
                                   labels[:batch_size_t2i, max_seq_length + 1:].contiguous().view(-1), ignore_index=-100)
    s_lm = batch_size_t2i
    e_lm = s_lm + batch_size_lm
    e_mmu = e_lm + batch_size_mmu
    e_r2i = e_mmu + batch_size_r2i
    masked = input_ids == mask_token_id

    def group(a, b, p_mask):
        m = masked[a:b]
        return F.cross_entropy(logits[a:b][m].contiguous().view(-1, V), labels[a:b][m].contiguous().view(-1),
                               ignore_index=-100, reduction="none") / p_mask[m], m

    loss_lm, m_lm = group(s_lm, e_lm, p_mask_lm)
    loss_lm = loss_lm.sum() / (logits[s_lm:e_lm].shape[0] * logits[s_lm:e_lm].shape[1])
    loss_lm = torch.sum(loss_lm / answer_lengths_lm[m_lm]) / logits[s_lm:e_lm].shape[0]
    loss_mmu, m_mmu = group(e_lm, e_mmu, p_mask_mmu)
    loss_mmu = torch.sum(loss_mmu / answer_lengths[m_mmu]) / logits[e_lm:e_mmu].shape[0]
    loss_r2i, m_r2i = group(e_mmu, e_r2i, p_mask_r2i)
    loss_r2i = torch.sum(loss_r2i / answer_lengths_r2i[m_r2i]) / logits[e_mmu:e_r2i].shape[0]
    return logits, loss_t2i, loss_lm, loss_mmu, loss_r2i


def forward_t2i(logits_fn, input_ids, labels, batch_size_t2i=0, max_seq_length=128, t2i_masks=None):
    """/root/reference/models/modeling_mmada.py:359-385."""
    logits = logits_fn(input_ids)
    V = logits.shape[-1]
    return F.cross_entropy(logits[:batch_size_t2i, max_seq_length + 1:].contiguous().view(-1, V),
                           labels[:batch_size_t2i, max_seq_length + 1:].contiguous().view(-1), ignore_index=-100)


def make_batch(B_t2i: int, B_lm: int, B_mmu: int, L: int, max_seq_length: int, seed: int, mask_token_id: int = 126336,
               text_vocab: int = 126349, codebook: int = 8192):
    """Synthetic mixed batch in the layout the training scripts build (train_mmada.py:520-600): t2i rows = text prefix of
    ``max_seq_length + 1`` positions then image tokens, some of them masked (labels elsewhere -100); lm / mmu rows = prompt
    then a partly masked answer; p_mask / answer lengths per position."""
    g = torch.Generator().manual_seed(seed)
    B = B_t2i + B_lm + B_mmu
    ids = torch.randint(0, 126000, (B, L), generator=g)
    labels = torch.full((B, L), -100, dtype=torch.long)
    p_lm = torch.rand(B_lm, L, generator=g) * 0.9 + 0.05
    p_mmu = torch.rand(B_mmu, L, generator=g) * 0.9 + 0.05
    al_lm = torch.zeros(B_lm, L, dtype=torch.long)
    al_mmu = torch.zeros(B_mmu, L, dtype=torch.long)
    t2i_masks = torch.ones(B_t2i, L, dtype=torch.long)
    for b in range(B_t2i):
        img = torch.randint(text_vocab, text_vocab + codebook, (L - max_seq_length - 1,), generator=g)
        m = torch.rand(L - max_seq_length - 1, generator=g) < 0.6
        ids[b, max_seq_length + 1:] = torch.where(m, torch.tensor(mask_token_id), img)
        labels[b, max_seq_length + 1:] = torch.where(m, img, torch.tensor(-100))
        t2i_masks[b, : 3 + b] = 0                                             # left padding of the text prefix
    for j, (n, p, al) in enumerate(((B_lm, p_lm, al_lm), (B_mmu, p_mmu, al_mmu))):
        for i in range(n):
            b = B_t2i + (0 if j == 0 else B_lm) + i
            start = 10 + 3 * i + 5 * j
            length = L - start - 2 * i
            tgt = torch.randint(0, 126000, (length,), generator=g)
            m = torch.rand(length, generator=g) < p[i, start:start + length]
            m[0] = True                                                       # at least one masked answer token
            ids[b, start:start + length] = torch.where(m, torch.tensor(mask_token_id), tgt)
            labels[b, start:start + length] = torch.where(m, tgt, torch.tensor(-100))
            al[i, :] = length
    return dict(input_ids=ids, labels=labels, p_mask_lm=p_lm, p_mask_mmu=p_mmu, answer_lengths=al_mmu,
                answer_lengths_lm=al_lm, t2i_masks=t2i_masks)

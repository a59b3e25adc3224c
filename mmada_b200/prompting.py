"""Token layout of the generation tasks for pre-tokenised text — host-side mirror of
``UniversalPrompting.t2i_gen_prompt`` / ``mmu_gen_prompt``
(/root/reference/training/prompting_utils.py:200-233, 379-425) with the reserved token ids of
:17-33.  The tokenizer itself is out of scope (no vocabulary offline); callers pass id lists.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch

RESERVED = {"<|soi|>": 126084, "<|eoi|>": 126085, "<|sov|>": 126086, "<|eov|>": 126087, "<|t2i|>": 126088,
            "<|mmu|>": 126089, "<|t2v|>": 126090, "<|v2v|>": 126091, "<|lvg|>": 126092, "[iPAD]": 126093,
            "<|r2i|>": 126094, "<|t2m|>": 126095, "<|som|>": 126096, "<|eom|>": 126097}
BOS, EOS = 126080, 126081
TEXT_VOCAB = 126349          # len(tokenizer) of the released checkpoints (reference app.py:396)
MASK_ID = 126336


def t2i_gen_prompt(text_ids: Sequence[Sequence[int]], image_ids: torch.Tensor, max_text_len: int = 512,
                   pad_id: int = RESERVED["[iPAD]"], bos: int = BOS, eos: int = EOS) -> Tuple[torch.Tensor, torch.Tensor]:
    """[pad.. <|t2i|> bos text eos] (max_text_len ids, left padded) <|soi|> image_ids <|eoi|>.
    Returns (sequence_ids, attention_mask), int64 (B, max_text_len + N + 2)."""
    n = image_ids.shape[-1]
    seqs, masks = [], []
    for i, t in enumerate(text_ids):
        t = list(t)
        if len(t) == 0:
            t = [bos]
        elif t[0] != bos:
            t = [bos] + t
        ids = [RESERVED["<|t2i|>"]] + t + [eos]
        if max_text_len >= len(ids):
            m = [0] * (max_text_len - len(ids)) + [1] * (len(ids) + n + 2)
            ids = [pad_id] * (max_text_len - len(ids)) + ids
        else:
            ids = ids[:max_text_len - 1] + [eos]
            m = [1] * (len(ids) + n + 2)
        seqs.append(torch.cat([torch.tensor(ids, dtype=torch.int64), torch.tensor([RESERVED["<|soi|>"]]),
                               image_ids[i].cpu().to(torch.int64), torch.tensor([RESERVED["<|eoi|>"]])]))
        masks.append(torch.tensor(m, dtype=torch.int64))
    return torch.stack(seqs), torch.stack(masks)


def mmu_prompt(image_code_ids: torch.Tensor, text_ids: Sequence[Sequence[int]], text_vocab: int = TEXT_VOCAB) -> List[torch.Tensor]:
    """<|mmu|> <|soi|> (codes + text_vocab) <|eoi|> text — the sequence inference_mmu.py:87-102 feeds to
    ``mmu_generate``.  Rows may differ in length; returned as a list."""
    out = []
    for i, t in enumerate(text_ids):
        out.append(torch.cat([torch.tensor([RESERVED["<|mmu|>"], RESERVED["<|soi|>"]]),
                              image_code_ids[i].cpu().to(torch.int64) + text_vocab,
                              torch.tensor([RESERVED["<|eoi|>"]]), torch.tensor(list(t), dtype=torch.int64)]))
    return out


def synthetic_t2i_batch(batch: int, max_text_len: int = 513, n_img: int = 1024, seed: int = 0, mask_id: int = MASK_ID):
    """Random pre-tokenised prompts (4..64 text ids each) and their empty-text unconditional twins,
    laid out by ``t2i_gen_prompt``.  Returns (cond_ids, uncond_ids, cond_mask, uncond_mask)."""
    g = torch.Generator().manual_seed(seed)
    img = torch.full((batch, n_img), mask_id, dtype=torch.int64)
    texts = []
    for _ in range(batch):
        t = int(torch.randint(4, min(64, max_text_len - 3) + 1, (1,), generator=g))
        texts.append(torch.randint(0, 126000, (t,), generator=g).tolist())
    cond, cm = t2i_gen_prompt(texts, img, max_text_len)
    unc, um = t2i_gen_prompt([[] for _ in range(batch)], img, max_text_len)
    return cond, unc, cm, um


class UniPromptingLike:
    """What the generate methods read from ``uni_prompting``: ``len(uni_prompting.text_tokenizer)``."""

    class _Tok:
        def __init__(self, n):
            self._n = n

        def __len__(self):
            return self._n

    def __init__(self, text_vocab: int = TEXT_VOCAB):
        self.text_tokenizer = self._Tok(text_vocab)

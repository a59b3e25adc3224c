"""Token layout of the generation tasks — mirror of the reference's ``UniversalPrompting``
(/root/reference/training/prompting_utils.py: ``t2i_gen_prompt`` :200-233, ``mmu_gen_prompt`` :379-425, ``t2m_prompt``
:87-144, ``__call__`` :482-540, reserved token ids :17-33) and of the inline MMU layout of inference_mmu.py:93-100.

``UniversalPrompting`` below assembles the (B, L) ids and masks ON THE DEVICE (csrc/prompting.cu) from ragged
pre-tokenised text; the tokenizer itself is out of scope (no vocabulary offline): ``__call__`` uses the tokenizer object
it is given, the ``*_prompt`` methods take id lists like the reference's.  The module-level functions are host-side
helpers for synthetic batches (benchmarks, tests).
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


class UniversalPrompting:
    """Device-side mirror of the reference class for the generation tasks (``use_reserved_token=True`` ids).  Same
    constructor arguments that matter for inference; ``text_tokenizer`` may be None when only id lists are passed
    (``bos`` / ``eos`` / ``end_header`` ids are then the defaults of the released checkpoints)."""

    def __init__(self, text_tokenizer=None, max_text_len: int = 8000, device="cuda", text_vocab: int = TEXT_VOCAB,
                 bos_token_id: int = BOS, eos_token_id: int = EOS, end_header_id: int = 126347, **_ignored):
        self.text_tokenizer = text_tokenizer if text_tokenizer is not None else UniPromptingLike._Tok(text_vocab)
        self.device = torch.device(device)
        self.max_text_len = max_text_len + 1            # the reference adds the task token's slot (:79)
        self.pad_id = RESERVED["[iPAD]"]
        self.bos = int(getattr(text_tokenizer, "bos_token_id", bos_token_id) or bos_token_id)
        self.eos = int(getattr(text_tokenizer, "eos_token_id", eos_token_id) or eos_token_id)
        self.end_header = int(end_header_id)
        self.sptids_dict = {k: torch.tensor([v]) for k, v in RESERVED.items()}
        self.sptids_dict["<|sot|>"] = torch.tensor([self.bos])
        self.sptids_dict["<|eot|>"] = torch.tensor([self.eos])
        self.sptids_dict["<|end_header_id|>"] = torch.tensor([self.end_header])

    def _ragged(self, text_ids):
        """list of id lists -> (flat int64, offsets int64 [B+1]) on the device: one host->device copy each."""
        off = [0]
        for t in text_ids:
            off.append(off[-1] + len(t))
        flat = torch.tensor([x for t in text_ids for x in t] or [0], dtype=torch.int64)
        return flat.to(self.device, non_blocking=True), torch.tensor(off, dtype=torch.int64).to(self.device, non_blocking=True)

    def _prefix(self, text_ids, body, task, open_tok, close_tok):
        from . import ops
        flat, off = self._ragged(text_ids)
        body = body.to(self.device, torch.int64)
        return ops.build_prompts(flat, off, body if body.stride(1) == 1 else body.contiguous(), self.max_text_len, 0,
                                 RESERVED[task], self.bos, self.eos, self.pad_id, RESERVED[open_tok], RESERVED[close_tok])

    def t2i_gen_prompt(self, text_ids, image_ids):
        """(:200-233) -> (sequence_ids, attention_mask), int64 (B, max_text_len + 1 + N + 2) on the device."""
        return self._prefix(text_ids, image_ids, "<|t2i|>", "<|soi|>", "<|eoi|>")

    def t2m_gen_prompt(self, text_ids, motion_ids):
        """``t2m_prompt`` (:87-144) as used for generation: no conditional drop-out, no labels."""
        return self._prefix(text_ids, motion_ids, "<|t2m|>", "<|som|>", "<|eom|>")

    def mmu_gen_prompt(self, image_ids, text_ids):
        """(:379-425) -> (sequence_ids (B, 3 + N + max_text_len), prompt_length (B,)).  The reference returns, per row,
        the mask ``[1] * prompt_length + [0] * max(0, max_text_len - prompt_length)`` (its second count is taken against
        the text part only); the prompt length is what that mask encodes."""
        from . import ops
        flat, off = self._ragged(text_ids)
        body = image_ids.to(self.device, torch.int64)
        return ops.build_prompts(flat, off, body if body.stride(1) == 1 else body.contiguous(), self.max_text_len - 1, 1,
                                 RESERVED["<|mmu|>"], self.bos, self.eos, self.pad_id, RESERVED["<|soi|>"],
                                 RESERVED["<|eoi|>"], self.end_header)

    def mmu_input_ids(self, image_tokens: torch.Tensor, text_ids: Sequence[int]) -> torch.Tensor:
        """inference_mmu.py:93-100: <|mmu|> <|soi|> image_tokens <|eoi|> <|sot|> text, one question for every row of
        ``image_tokens`` (already offset by ``len(text_tokenizer)``); assembled on the device."""
        dev = self.device
        B = image_tokens.shape[0]
        head = torch.tensor([RESERVED["<|mmu|>"], RESERVED["<|soi|>"]], dtype=torch.int64, device=dev).expand(B, 2)
        tail = torch.tensor([RESERVED["<|eoi|>"], self.bos] + list(text_ids), dtype=torch.int64, device=dev).expand(B, -1)
        return torch.cat([head, image_tokens.to(dev, torch.int64), tail], dim=1)

    def __call__(self, input, task, padding=True, config=None):
        """``uni_prompting((prompts, image_tokens), 't2i_gen')`` as inference_t2i.py:92-94 calls it (needs a tokenizer)."""
        if task == "t2i_gen":
            text_ids = self.text_tokenizer(input[0])["input_ids"]
            return self.t2i_gen_prompt(text_ids, input[1])
        if task == "t2m_gen":
            text_ids = self.text_tokenizer(input[0])["input_ids"]
            return self.t2m_gen_prompt(text_ids, input[1])
        raise NotImplementedError(f"task {task!r}: only the generation layouts are built here (training prompts are out of scope)")


class UniPromptingLike:
    """What the generate methods read from ``uni_prompting``: ``len(uni_prompting.text_tokenizer)``."""

    class _Tok:
        def __init__(self, n):
            self._n = n

        def __len__(self):
            return self._n

    def __init__(self, text_vocab: int = TEXT_VOCAB):
        self.text_tokenizer = self._Tok(text_vocab)

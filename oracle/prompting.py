"""Oracle: sequence layout of the generation tasks, restated on Python lists.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows /root/reference/training/prompting_utils.py:
  * UniversalPrompting.t2i_gen_prompt  :200-233   [pad.. <|t2i|> bos text eos] <|soi|> image <|eoi|>, attention mask
  * UniversalPrompting.mmu_gen_prompt  :379-425   <|mmu|> <|soi|> image <|eoi|> [bos text eos, eos padding], prompt length
  * UniversalPrompting.t2m_prompt      :87-144    [pad.. <|t2m|> bos text eos] <|som|> motion <|eom|> (cond. drop-out off)
  * reserved ids :17-33; ``self.max_text_len = max_text_len + 1`` :79; pad id = [iPAD] :80
and the inline layout of inference_mmu.py:93-100 (<|mmu|> <|soi|> image <|eoi|> <|sot|> text).
Pinned against the real class by oracle/make_goldens.py (stub tokenizer; tests/golden/prompting.npz).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

RESERVED = {"<|soi|>": 126084, "<|eoi|>": 126085, "<|sov|>": 126086, "<|eov|>": 126087, "<|t2i|>": 126088,
            "<|mmu|>": 126089, "<|t2v|>": 126090, "<|v2v|>": 126091, "<|lvg|>": 126092, "[iPAD]": 126093,
            "<|r2i|>": 126094, "<|t2m|>": 126095, "<|som|>": 126096, "<|eom|>": 126097}
BOS, EOS = 126080, 126081
END_HEADER = 126347          # stand-in for the tokenizer's <|end_header_id|> in the synthetic cases


def _with_bos(t: Sequence[int], bos: int) -> List[int]:
    t = list(t)
    if len(t) == 0:
        return [bos]
    return t if t[0] == bos else [bos] + t


def prefix_layout(text: Sequence[int], body: Sequence[int], max_text_len: int, task: int, open_tok: int, close_tok: int,
                  pad: int = RESERVED["[iPAD]"], bos: int = BOS, eos: int = EOS) -> Tuple[List[int], List[int]]:
    """t2i_gen_prompt / t2m_prompt for one row.  ``max_text_len`` is the constructor argument (the class adds 1)."""
    P = max_text_len + 1
    ids = [task] + _with_bos(text, bos) + [eos]
    if P >= len(ids):
        mask = [0] * (P - len(ids)) + [1] * (len(ids) + len(body) + 2)
        ids = [pad] * (P - len(ids)) + ids
    else:
        ids = ids[:P - 1] + [eos]
        mask = [1] * (len(ids) + len(body) + 2)
    return ids + [open_tok] + list(body) + [close_tok], mask


def mmu_gen_layout(text: Sequence[int], image: Sequence[int], max_text_len: int, end_header: int = END_HEADER,
                   bos: int = BOS, eos: int = EOS) -> Tuple[List[int], int]:
    """mmu_gen_prompt for one row: (sequence ids, prompt_length)."""
    temp = _with_bos(text, bos) + [eos]
    if max_text_len >= len(temp):
        temp = temp + [eos] * (max_text_len - len(temp))
    else:
        temp = temp[:max_text_len - 1] + [eos]
    seq = [RESERVED["<|mmu|>"], RESERVED["<|soi|>"]] + list(image) + [RESERVED["<|eoi|>"]] + temp
    pos = -1
    for i in range(len(temp) - 1, -1, -1):
        if temp[i] == end_header:
            pos = i
            break
    prompt_length = len(seq) - len(temp) + (pos + 1 if pos != -1 else 0)
    return seq, prompt_length


def mmu_gen_mask(prompt_length: int, n_text_slots: int) -> List[int]:
    """The prompt mask mmu_gen_prompt returns for a row (:416-419): ``[1] * prompt_length + [0] * predict_length`` with
    ``predict_length = len(text part) - prompt_length`` — counted against the TEXT part only, so with an image in front
    it is shorter than the sequence (and just ``prompt_length`` ones once that exceeds the text slots).  Kept as is."""
    return [1] * prompt_length + [0] * max(0, n_text_slots - prompt_length)


def mmu_inference_layout(text: Sequence[int], image_tokens: Sequence[int], bos: int = BOS) -> List[int]:
    """inference_mmu.py:93-100 for one row (image_tokens already offset by len(tokenizer))."""
    return [RESERVED["<|mmu|>"], RESERVED["<|soi|>"]] + list(image_tokens) + [RESERVED["<|eoi|>"], bos] + list(text)

"""Mask schedules and MaskGIT confidence re-masking — mirror of the reference's models/sampling.py
(/root/reference/models/sampling.py:10-16, 31-36, 39-77), same names and argument meaning.

The schedules are scalar host math (the reference evaluates them on 0-d CPU tensors,
modeling_mmada.py:186-187) and stay plain torch; ``mask_by_random_topk`` runs the CUDA kernel.
"""
from __future__ import annotations

import math
from functools import partial
from typing import Optional

import torch

from . import ops


def cosine_schedule(t):
    return torch.cos(t * math.pi * 0.5)


def linear_schedule(t):
    mask_ratio = 1 - t
    return mask_ratio.clamp(min=1e-6, max=1.0)


def pow(t, method):
    exponent = float(method.replace("pow", ""))
    mask_ratio = 1.0 - t ** exponent
    return mask_ratio.clamp(min=1e-6, max=1.0)


def sigmoid_schedule(t, start=-3, end=3, tau=1.0, clip_min=1e-6):
    v_start = torch.sigmoid(torch.tensor(start / tau))
    v_end = torch.sigmoid(torch.tensor(end / tau))
    output = torch.sigmoid((t * (end - start) + start) / tau)
    output = (v_end - output) / (v_end - v_start)
    return torch.clip(output, clip_min, 1.0)


def get_mask_schedule(method, **schedule_kwargs):
    if method == "cosine":
        return cosine_schedule
    elif method == "linear":
        return linear_schedule
    elif "pow" in method:
        return partial(pow, method=method)
    elif method == "sigmoid":
        return partial(sigmoid_schedule, **schedule_kwargs)
    else:
        raise ValueError("Unknown schedule method: {}".format(method))


def mask_by_random_topk(mask_len, probs, temperature=1.0, generator: Optional[torch.Generator] = None):
    """masking = confidence < sorted(confidence)[mask_len], confidence = log(probs) + T * gumbel.
    The uniform noise is drawn exactly like the reference draws it (``zeros_like(probs).uniform_(0, 1,
    generator=generator)``), so a given generator state yields the reference's own noise."""
    noise = torch.zeros_like(probs).uniform_(0, 1, generator=generator)
    return ops.mask_by_random_topk(mask_len, probs.float(), noise.float(), float(temperature))

"""mmada_b200 — B200-native (sm_100a) implementation of MMaDA's masked-diffusion denoising path.

Public surface mirrors the reference's ``models`` package for this path
(/root/reference/models/__init__.py): ``MMadaModelLM``, ``MMadaConfig``, the schedules and
``mask_by_random_topk`` of ``models/sampling.py``, ``MAGVITv2`` (decode_code / get_code); plus ``generate``
(reference generate.py:43) and ``HumanVQVAE.forward_decoder`` (motion_vqvae/models/vqvae.py:74-81,115-117).  Everything executes hand-written CUDA kernels from libmmada_b200.so through a
C ABI (include/mmada_b200.h); there is no CPU or PyTorch fallback — importing works without a GPU,
calling raises.
"""
from .generate import generate, get_num_transfer_tokens  # noqa: F401
from .modeling_llada import LLaDAConfig, LLaDAModelLM, interleave_gate_up  # noqa: F401
from .modeling_mmada import MMadaConfig, MMadaModelLM  # noqa: F401
from .modeling_magvitv2 import MAGVITv2  # noqa: F401
from .motion_vqvae import HumanVQVAE  # noqa: F401
from .sampling import (cosine_schedule, get_mask_schedule, linear_schedule, mask_by_random_topk,  # noqa: F401
                       sigmoid_schedule)

__version__ = "0.1.0"

"""Checkpoint entry: ``from_pretrained(dir)`` for the LLM (Hugging Face layout) and the VQ model.

Mirror of what the reference gets from ``MMadaModelLM.from_pretrained(path, torch_dtype=bf16)`` /
``MAGVITv2.from_pretrained(path)`` (/root/reference/inference_t2i.py:66-70; file names: transformers' ``config.json`` +
``model.safetensors`` | ``model-00001-of-0000N.safetensors`` + ``model.safetensors.index.json`` | ``pytorch_model.bin``,
and for the VQ model ``config.json`` + ``pytorch_model.safetensors`` | ``pytorch_model.bin``,
models/modeling_utils.py:47-49).  Local directories only (no hub access offline).  Tensors are streamed one at a time
from the shards into the layouts the kernels use: the 8B checkpoint never sits in host memory as a whole.
"""
from __future__ import annotations

import json
import os
from collections.abc import Mapping
from typing import Dict, List, Optional

import torch


class ShardedCheckpoint(Mapping):
    """Lazy key -> tensor view over the weight files of a checkpoint directory."""

    def __init__(self, path: str, names=("model", "pytorch_model", "diffusion_pytorch_model")):
        self.path = path
        self._where: Dict[str, str] = {}          # key -> file
        self._open = {}                           # file -> safe_open handle / loaded .bin dict
        files: List[str] = []
        for stem in names:
            idx = os.path.join(path, stem + ".safetensors.index.json")
            if os.path.isfile(idx):
                wm = json.load(open(idx))["weight_map"]
                self._where = {k: os.path.join(path, f) for k, f in wm.items()}
                return
        for stem in names:
            f = os.path.join(path, stem + ".safetensors")
            if os.path.isfile(f):
                files = [f]
                break
        if not files:
            files = sorted(os.path.join(path, f) for f in os.listdir(path) if f.endswith(".safetensors"))
        if files:
            from safetensors import safe_open
            for f in files:
                with safe_open(f, framework="pt") as h:
                    for k in h.keys():
                        self._where[k] = f
            return
        for stem in names:
            f = os.path.join(path, stem + ".bin")
            if os.path.isfile(f):
                sd = torch.load(f, map_location="cpu", weights_only=True)
                self._open[f] = sd
                self._where = {k: f for k in sd}
                return
        raise FileNotFoundError(f"no *.safetensors / *.bin weights under {path}")

    def __getitem__(self, key: str) -> torch.Tensor:
        f = self._where[key]
        h = self._open.get(f)
        if h is None:
            from safetensors import safe_open
            h = self._open[f] = safe_open(f, framework="pt")
        return h[key] if isinstance(h, dict) else h.get_tensor(key)

    def __iter__(self):
        return iter(self._where)

    def __len__(self):
        return len(self._where)


def read_config(path: str) -> dict:
    f = os.path.join(path, "config.json")
    if not os.path.isfile(f):
        raise FileNotFoundError(f"{f} not found")
    return json.load(open(f))


def save_pretrained_llm(path: str, config: dict, state_dict: Dict[str, torch.Tensor], max_shard_bytes: Optional[int] = None):
    """Write ``config.json`` + (sharded) safetensors with the reference's key names — what tests use to make a synthetic
    checkpoint (there is no network for real ones)."""
    from safetensors.torch import save_file
    os.makedirs(path, exist_ok=True)
    json.dump(config, open(os.path.join(path, "config.json"), "w"), indent=1)
    if not max_shard_bytes:
        save_file({k: v.contiguous() for k, v in state_dict.items()}, os.path.join(path, "model.safetensors"))
        return
    shards, cur, size = [], {}, 0
    for k, v in state_dict.items():
        nb = v.numel() * v.element_size()
        if cur and size + nb > max_shard_bytes:
            shards.append(cur)
            cur, size = {}, 0
        cur[k] = v.contiguous()
        size += nb
    shards.append(cur)
    wm = {}
    for i, sh in enumerate(shards):
        name = f"model-{i + 1:05d}-of-{len(shards):05d}.safetensors"
        save_file(sh, os.path.join(path, name))
        wm.update({k: name for k in sh})
    json.dump({"metadata": {}, "weight_map": wm}, open(os.path.join(path, "model.safetensors.index.json"), "w"))

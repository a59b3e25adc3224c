"""Checkpoint entry (SURVEY.md 8(b) last row): ``from_pretrained(dir)`` over a synthetic checkpoint written with the
reference's key names (Appendix D) — single-file and sharded safetensors — and the configuration guard."""
import dataclasses
import json
import os

import pytest
import torch


def _hf_config(cfg):
    """A config.json like the released MMaDA checkpoints' (reference ModelConfig fields + MMadaConfig's own keys)."""
    return dict(cfg, model_type="mmada", n_kv_heads=cfg["n_heads"], block_type="llama", activation_type="silu",
                layer_norm_type="rms", rope=True, alibi=False, include_bias=False, weight_tying=False, scale_logits=False,
                input_emb_norm=False, attention_layer_norm=False, embedding_size=cfg["vocab_size"], rope_full_precision=True,
                llm_vocab_size=126464, codebook_size=8192, num_vq_tokens=256, num_new_special_tokens=0,
                new_vocab_size=cfg["vocab_size"], torch_dtype="bfloat16")


def test_sharded_checkpoint_reader_and_config_guard(tmp_path):
    from mmada_b200 import LLaDAConfig, MMadaConfig
    from mmada_b200.checkpoint import ShardedCheckpoint, read_config, save_pretrained_llm
    from oracle import weights as W
    cfg = dict(W.TINY, vocab_size=1024)
    sd = W.make_llada_weights(cfg, 3)
    one, many = str(tmp_path / "one"), str(tmp_path / "many")
    save_pretrained_llm(one, _hf_config(cfg), sd)
    save_pretrained_llm(many, _hf_config(cfg), sd, max_shard_bytes=600_000)
    assert len([f for f in os.listdir(many) if f.endswith(".safetensors")]) > 2
    for path in (one, many):
        ck = ShardedCheckpoint(path)
        assert set(ck.keys()) == set(sd.keys())
        for k in list(sd)[::5]:
            assert torch.equal(ck[k], sd[k])
        c = MMadaConfig.from_dict(read_config(path))
        assert (c.d_model, c.n_heads, c.n_layers, c.vocab_rows) == (cfg["d_model"], cfg["n_heads"], cfg["n_layers"], 1024)
    # options these kernels do not implement are refused instead of loading to different logits
    for bad in (dict(n_kv_heads=1), dict(include_bias=True), dict(weight_tying=True), dict(scale_logits=True),
                dict(attention_layer_norm=True), dict(block_type="sequential"), dict(rope=False), dict(alibi=True),
                dict(activation_type="swiglu"), dict(layer_norm_type="default")):
        with pytest.raises(ValueError):
            LLaDAConfig.from_dict(dict(_hf_config(cfg), **bad))
    # the reference's own enum members are accepted (configuration_llada.py:80-127)
    class _E(str):
        @property
        def value(self):
            return str(self)
    LLaDAConfig.from_dict(dict(_hf_config(cfg), block_type=_E("llama")))


def test_state_dict_key_guard():
    from mmada_b200 import LLaDAModelLM
    with pytest.raises(ValueError):
        LLaDAModelLM.check_state_dict_keys(["model.transformer.blocks.0.q_proj.bias"])
    with pytest.raises(ValueError):
        LLaDAModelLM.check_state_dict_keys(["model.transformer.blocks.0.q_norm.weight"])
    LLaDAModelLM.check_state_dict_keys(["model.transformer.blocks.0.q_proj.weight", "model.transformer.wte.weight"])


@pytest.mark.gpu
def test_from_pretrained_matches_load_state_dict(tmp_path):
    from mmada_b200 import MAGVITv2, MMadaModelLM, MMadaConfig
    from mmada_b200.checkpoint import save_pretrained_llm
    from oracle import weights as W
    from safetensors.torch import save_file
    cfg = dict(W.TINY128)
    sd = W.make_llada_weights(cfg, 4)
    path = str(tmp_path / "llm")
    save_pretrained_llm(path, _hf_config(cfg), {k: v.to(torch.bfloat16) for k, v in sd.items()}, max_shard_bytes=40_000_000)
    a = MMadaModelLM.from_pretrained(path, torch_dtype=torch.bfloat16)
    b = MMadaModelLM(MMadaConfig.from_dict(cfg)).load_state_dict({k: v.to(torch.bfloat16) for k, v in sd.items()})
    ids = torch.randint(0, 126000, (2, 70), generator=torch.Generator().manual_seed(1)).cuda()
    assert torch.equal(a(ids).logits, b(ids).logits)
    assert a.hf_config["codebook_size"] == 8192
    # a grouped-query checkpoint (k_proj with fewer rows) is refused by shape, whatever its config says
    bad = dict(sd)
    bad["model.transformer.blocks.0.k_proj.weight"] = bad["model.transformer.blocks.0.k_proj.weight"][:128]
    with pytest.raises(ValueError):
        MMadaModelLM(MMadaConfig.from_dict(cfg)).load_state_dict(bad)
    # VQ model: config.json + pytorch_model.safetensors (models/modeling_utils.py:47-49)
    vq_dir = tmp_path / "vq"
    vq_dir.mkdir()
    vsd = W.make_vq_decoder_weights(0)
    json.dump({"_class_name": "MAGVITv2"}, open(vq_dir / "config.json", "w"))
    save_file({k: v.contiguous() for k, v in vsd.items()}, str(vq_dir / "pytorch_model.safetensors"))
    codes = torch.randint(0, 8192, (1, 256), generator=torch.Generator().manual_seed(2)).cuda()
    assert torch.equal(MAGVITv2.from_pretrained(str(vq_dir)).decode_code(codes), MAGVITv2().load_state_dict(vsd).decode_code(codes))

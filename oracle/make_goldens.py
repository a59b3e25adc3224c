"""Pin the oracle against the REAL reference and write tests/golden/*.npz.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Run in the build container, where /root/reference
exists:   python -m oracle.make_goldens
For every case the real reference (imported by oracle/ref_bootstrap.py, unmodified) is executed on
the synthetic weights of oracle/weights.py; the restatement in oracle/{llada,denoise,magvit}.py
is asserted bit-identical to it; and the REFERENCE's outputs are stored as the golden vectors.
The fixtures are small (ids, masks, sub-sampled logits, checksums); big inputs (weights, noise)
are regenerated deterministically from seeds at test time.
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

from . import denoise, llada, magvit, ref_bootstrap as rb, training, weights as W

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _save(name, **arrs):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: (v.detach().cpu().numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in arrs.items()})
    print(f"  wrote {path} ({os.path.getsize(path) / 1024:.1f} KiB)")


def t2i_case(name, cfg, B, P, N, steps, guidance, wseed, pseed, gseed):
    print(f"[t2i] {name}")
    sd = W.make_llada_weights(cfg, wseed)
    model = rb.build_model(cfg, sd)
    cond, unc, am, uam = W.make_t2i_prompts(B, P, N, pseed)
    # --- the real reference
    ids_ref = cond.clone()
    g = torch.Generator().manual_seed(gseed)
    t0 = time.time()
    out_ref = model.t2i_generate(input_ids=ids_ref, uncond_input_ids=unc.clone(), attention_mask=am,
                                 uncond_attention_mask=uam, guidance_scale=guidance, timesteps=steps,
                                 seq_len=N, resolution=P - 1, generator=g,
                                 uni_prompting=rb.UniPromptingStub(W.TEXT_VOCAB))
    print(f"  reference t2i_generate: {time.time() - t0:.1f}s")
    # --- the restatement, full-vocab logits, same generator seed
    ids_or = cond.clone()
    trace = []
    out_or = denoise.t2i_generate(lambda x: llada.forward_logits(x, sd, cfg), ids_or, unc.clone(),
                                  guidance_scale=guidance, timesteps=steps, seq_len=N, resolution=P - 1,
                                  generator=torch.Generator().manual_seed(gseed), trace=trace)
    assert torch.equal(out_ref, out_or), "restatement != reference (sampled_ids)"
    assert torch.equal(ids_ref, ids_or), "restatement != reference (mutated input_ids)"
    # --- restricted lm_head (image rows x codebook columns) must give the same decisions
    ids_sl = cond.clone()
    rows, cols = slice(-(N + 1), -1), slice(W.TEXT_VOCAB, W.TEXT_VOCAB + W.CODEBOOK)
    out_sl = denoise.t2i_generate(lambda x: llada.forward_logits(x, sd, cfg, rows, cols), ids_sl, unc.clone(),
                                  guidance_scale=guidance, timesteps=steps, seq_len=N, resolution=P - 1,
                                  generator=torch.Generator().manual_seed(gseed), sliced_logits=True)
    same = torch.equal(out_sl, out_ref) and torch.equal(ids_sl, ids_ref)
    print(f"  restricted lm_head reproduces reference ids: {same}")
    # the reference's own first-forward logits (for the tolerance-based logits parity test)
    with torch.no_grad():
        lg = model(torch.cat([cond, unc])).logits[:, rows, cols]
    _save(name,
          sampled_ids=out_ref, final_input_ids=ids_ref, cond_ids=cond, uncond_ids=unc,
          step_sampled=torch.stack([t["sampled_ids"] for t in trace]),
          step_masking=torch.stack([t["masking"] for t in trace]),
          step_mask_len=torch.stack([t["mask_len"] for t in trace]),
          step_sel=torch.stack([t["selected_probs"] for t in trace]),
          step_u=torch.stack([t["u"] for t in trace]),
          step_q_sum=np.array([float(t["q"].double().sum()) for t in trace]),
          step_temperature=np.array([t["temperature"] for t in trace]),
          step_mask_len_raw=np.array([t["mask_len_raw"] for t in trace]),
          step_cond_sub=torch.stack([t["cond"][:, ::16, ::64] for t in trace]),
          step_uncond_sub=torch.stack([t["uncond"][:, ::16, ::64] for t in trace]),
          first_logits_sub=lg[:, ::4, ::32], first_logits_absmax=lg.abs().max(), first_logits_std=lg.std(),
          restricted_same=np.array(same),
          meta=np.array([B, P, N, steps, wseed, pseed, gseed]), guidance=np.array(guidance))


def text_case(name, cfg, B, Lp, gen, block, steps, temperature, cfg_scale, wseed, seed, fast_eot=None):
    print(f"[text] {name}")
    sd = W.make_llada_weights(cfg, wseed)
    model = rb.build_model(cfg, sd)
    g = torch.Generator().manual_seed(seed)
    prompt = torch.randint(0, 126000, (B, Lp), generator=g)
    gen_fn = rb.load_generate_fn()
    torch.manual_seed(seed)
    x_ref = gen_fn(model, prompt, steps=steps, gen_length=gen, block_length=block, temperature=temperature,
                   cfg_scale=cfg_scale, remasking="low_confidence")
    torch.manual_seed(seed)
    x_mmu = model.mmu_generate(idx=prompt, max_new_tokens=gen, steps=steps, block_length=block,
                               temperature=temperature, cfg_scale=cfg_scale)
    assert torch.equal(x_ref, x_mmu), "generate() != mmu_generate()"
    torch.manual_seed(seed)
    trace = []
    x_or = denoise.generate(lambda x: llada.forward_logits(x, sd, cfg), prompt, steps=steps, gen_length=gen,
                            block_length=block, temperature=temperature, cfg_scale=cfg_scale, trace=trace)
    assert torch.equal(x_ref, x_or), "restatement != reference (generate)"
    extra = {}
    if fast_eot is not None:
        eot = int(x_ref[0, Lp + block - 1]) if fast_eot == "hit" else 5
        torch.manual_seed(seed)
        x_fast = model.mmu_generate_fast(idx=prompt[:1], max_new_tokens=gen, steps=steps, block_length=block,
                                         temperature=temperature, cfg_scale=cfg_scale, eot_token=eot)
        torch.manual_seed(seed)
        x_fo = denoise.generate(lambda x: llada.forward_logits(x, sd, cfg), prompt[:1], steps=steps,
                                gen_length=gen, block_length=block, temperature=temperature,
                                cfg_scale=cfg_scale, eot_token=eot)
        assert torch.equal(x_fast, x_fo), "restatement != reference (mmu_generate_fast)"
        extra = dict(fast_x=x_fast, fast_eot=np.array(eot))
    _save(name, x=x_ref, prompt=prompt,
          step_x=torch.stack([t["x"] for t in trace]),
          step_transfer=torch.stack([t["transfer"] for t in trace]),
          meta=np.array([B, Lp, gen, block, steps, wseed, seed]),
          temperature=np.array(temperature), cfg_scale=np.array(cfg_scale), **extra)


def t2m_case(name, cfg, B, Lt, N, steps, wseed, seed, gseed):
    """t2m_generate of the reference's fork models/modelling_ours.py (BASELINE configs[4])."""
    import importlib
    print(f"[t2m] {name}")
    rb.modules()
    mo = importlib.import_module("models.modelling_ours")
    sd = W.make_llada_weights(cfg, wseed)
    base = rb.build_model(cfg, sd)
    # the fork's class has the same constructor; reuse the loaded reference model's weights
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        model = mo.MMadaModelLM(base.config, init_params=False).eval()
    model.load_state_dict(base.state_dict())
    g = torch.Generator().manual_seed(seed)
    som, eom = 126096, 126097
    text = torch.randint(0, 126000, (B, Lt), generator=g)
    ids = torch.cat([text, torch.full((B, 1), som), torch.full((B, N), 126336), torch.full((B, 1), eom)], 1)

    class UP(rb.UniPromptingStub):
        sptids_dict = {"<|som|>": torch.tensor([som]), "<|eom|>": torch.tensor([eom])}

    ids_ref = ids.clone()
    out_ref = model.t2m_generate(input_ids=ids_ref, timesteps=steps, seq_len=N, generator=torch.Generator().manual_seed(gseed),
                                 uni_prompting=UP(W.TEXT_VOCAB))
    ids_or = ids.clone()
    trace = []
    out_or = denoise.t2m_generate(lambda x: llada.forward_logits(x, sd, cfg), ids_or, timesteps=steps, seq_len=N,
                                  generator=torch.Generator().manual_seed(gseed), som_token=som, eom_token=eom, trace=trace)
    assert torch.equal(out_ref, out_or), "restatement != reference (t2m sampled_ids)"
    assert torch.equal(ids_ref, ids_or), "restatement != reference (t2m input_ids)"
    _save(name, ids=ids, sampled_ids=out_ref, final_input_ids=ids_ref,
          step_merged=torch.stack([t["merged"] for t in trace]),
          step_masking=torch.stack([t["masking"] for t in trace[:-1]]),
          meta=np.array([B, Lt, N, steps, wseed, seed, gseed]))


def sampling_case():
    print("[sampling]")
    _, _, sp, _ = rb.modules()
    tabs = {}
    for N in (64, 256, 1024):
        for T in (8, 15, 18):
            ref = [float((N * sp.cosine_schedule(torch.tensor(1.0 * (s + 1) / T))).floor()) for s in range(T)]
            assert ref == denoise.t2i_mask_len_schedule(N, T)
            tabs[f"cos_{N}_{T}"] = np.array(ref)
    assert list(tabs["cos_1024_15"][[9, 14]]) == [511.0, -1.0]
    g = torch.Generator().manual_seed(7)
    probs = torch.rand(4, 256, generator=g) ** 4
    probs[:, ::5] = torch.finfo(torch.float32).max
    probs[1, 10:40] = probs[1, 10]                      # ties at/around the cut-off
    mask_len = torch.tensor([[1.0], [20.0], [100.0], [200.0]])
    outs = []
    for T in (0.0, 0.3, 1.0):
        gr = torch.Generator().manual_seed(11)
        ref = sp.mask_by_random_topk(mask_len, probs, T, generator=gr)
        go = torch.Generator().manual_seed(11)
        assert torch.equal(ref, denoise.mask_by_random_topk(mask_len, probs, T, generator=go))
        outs.append(ref)
    u = torch.zeros_like(probs).uniform_(0, 1, generator=torch.Generator().manual_seed(11))
    for name in ("linear", "pow2", "sigmoid"):
        f = sp.get_mask_schedule(name)
        tabs["sched_" + name] = np.array([float(f(torch.tensor(t))) for t in np.linspace(0, 1, 11)])
    _save("sampling", probs=probs, mask_len=mask_len, u=u, masking=torch.stack(outs),
          temperatures=np.array([0.0, 0.3, 1.0]), **tabs)


def magvit_case():
    print("[magvit]")
    sd = W.make_vq_decoder_weights(0)
    vq = rb.build_vq(sd)
    # LFQ tables against the reference's buffers
    all_idx = torch.arange(8192).view(8, 1024)
    ref_bits = vq.quantize.get_codebook_entry(all_idx)
    assert np.array_equal(ref_bits.numpy(), magvit.lfq_indices_to_bits(all_idx.numpy()))
    ref_idx = vq.quantize.get_indices(ref_bits)
    assert np.array_equal(ref_idx.numpy(), magvit.lfq_bits_to_indices(ref_bits.numpy()))
    assert torch.equal(ref_idx.reshape(8, 1024), all_idx)
    g = torch.Generator().manual_seed(3)
    out = {}
    for tag, B, n in (("8x8", 2, 64), ("16x16", 1, 256)):
        idx = torch.randint(0, 8192, (B, n), generator=g)
        with torch.no_grad():
            ref = vq.decode_code(idx)
        taps = {}
        mine = magvit.decode_code(idx, sd, taps)
        assert torch.equal(ref, mine), "restatement != reference (decode_code)"
        out[f"idx_{tag}"] = idx
        out[f"pix_{tag}"] = ref if n == 64 else ref[:, :, ::2, ::2]
        out[f"pix_absmax_{tag}"] = ref.abs().max()
        out[f"tap_names_{tag}"] = np.array(list(taps.keys()))
        out[f"tap_mean_{tag}"] = np.array([float(v.double().mean()) for v in taps.values()])
        out[f"tap_std_{tag}"] = np.array([float(v.double().std()) for v in taps.values()])
    _save("magvit", bits_sample_idx=torch.tensor([0, 5, 4096, 8191]),
          bits_sample=vq.quantize.get_codebook_entry(torch.tensor([[0, 5, 4096, 8191]])), **out)


def magvit_encoder_case():
    """MAGVITv2.get_code: the real reference's encoder on synthetic weights, 256 x 256 input
    (16 x 16 tokens: the smallest grid the GPU convolution tiling accepts)."""
    print("[magvit encoder]")
    sd = W.make_vq_encoder_weights(0)
    vq = rb.build_vq(sd)
    g = torch.Generator().manual_seed(11)
    # smooth-ish image in [-1, 1]: low-resolution noise upsampled, plus fine noise
    base = torch.nn.functional.interpolate(torch.randn(1, 3, 16, 16, generator=g), size=(256, 256), mode="bilinear")
    px = (0.6 * base + 0.2 * torch.randn(1, 3, 256, 256, generator=g)).clamp(-1, 1)
    px = px.to(torch.float16).float()          # the fixture stores fp16: both sides get exactly these values
    with torch.no_grad():
        ref_z = vq.encoder(px)
        ref_codes = vq.get_code(px)
    mine_z = magvit.encoder_forward(px, sd)
    assert torch.equal(ref_z, mine_z), "restatement != reference (encoder)"
    assert torch.equal(ref_codes, magvit.get_code(px, sd)), "restatement != reference (get_code)"
    _save("magvit_encoder", pixels=px.to(torch.float16), latents=ref_z, codes=ref_codes,
          latent_absmax=ref_z.abs().max(), note=np.array("pixels are exactly representable in fp16: feed pixels.float()"))


def motion_case():
    """Motion VQ-VAE decode: the reference's own Decoder class on synthetic weights."""
    print("[motion decoder]")
    import sys
    from . import motion
    if rb.REF not in sys.path:
        sys.path.insert(0, rb.REF)
    from motion_vqvae.models.encdec import Decoder
    cfg = motion.MOTION
    sd = motion.make_motion_decoder_weights(0)
    dec = Decoder(cfg["n_feats"], cfg["code_dim"], cfg["down_t"], 2, cfg["width"], cfg["depth"], cfg["rate"], activation="relu", norm=None).eval()
    missing, unexpected = dec.load_state_dict({k[len("vqvae.decoder."):]: v for k, v in sd.items() if ".decoder." in k}, strict=True)
    ids = torch.randint(0, cfg["nb_code"], (1, 49), generator=torch.Generator().manual_seed(21))
    with torch.no_grad():
        # vqvae.py:74-81 around the reference's Decoder
        x_d = torch.nn.functional.embedding(ids, sd["vqvae.quantizer.codebook"]).view(1, -1, cfg["code_dim"]).permute(0, 2, 1).contiguous()
        ref = dec(x_d).permute(0, 2, 1)
    mine = motion.forward_decoder(ids, sd)
    assert torch.equal(ref, mine), "restatement != reference (motion decoder)"
    _save("motion_decoder", ids=ids, pose=ref, pose_absmax=ref.abs().max())


def logits_case(name, cfg, B, L, wseed, seed):
    print(f"[logits] {name}")
    sd = W.make_llada_weights(cfg, wseed)
    model = rb.build_model(cfg, sd)
    ids = torch.randint(0, 126000, (B, L), generator=torch.Generator().manual_seed(seed))
    ids[:, L // 2:] = cfg["mask_token_id"]
    with torch.no_grad():
        ref = model(ids).logits
        # Q1: the bias is never applied
        bias = torch.zeros(B, 1, L, L, dtype=torch.bool)
        assert torch.equal(ref, model(ids, attention_bias=bias).logits)
    mine = llada.forward_logits(ids, sd, cfg)
    assert torch.equal(ref, mine), "restatement != reference (logits)"
    cols = slice(W.TEXT_VOCAB, W.TEXT_VOCAB + W.CODEBOOK)
    _save(name, ids=ids, logits_img=ref[:, ::3, cols][:, :, ::16], logits_txt=ref[:, ::3, :126000:512],
          absmax=ref.abs().max(), std=ref.std(), meta=np.array([B, L, wseed, seed]))


def forward_process_case(name, cfg, B_t2i, B_lm, B_mmu, L, max_seq_length, wseed, seed):
    """MMadaModelLM.forward_process (modeling_mmada.py:213-276, forward values): the real method against
    oracle/training.py on a synthetic mixed t2i / lm / mmu batch."""
    print(f"[forward_process] {name}")
    sd = W.make_llada_weights(cfg, wseed)
    model = rb.build_model(cfg, sd)
    bt = training.make_batch(B_t2i, B_lm, B_mmu, L, max_seq_length, seed, cfg["mask_token_id"])
    kw = dict(batch_size_t2i=B_t2i, batch_size_lm=B_lm, batch_size_mmu=B_mmu, max_seq_length=max_seq_length,
              p_mask_lm=bt["p_mask_lm"], p_mask_mmu=bt["p_mask_mmu"], answer_lengths=bt["answer_lengths"],
              t2i_masks=bt["t2i_masks"], answer_lengths_lm=bt["answer_lengths_lm"])
    with torch.no_grad():
        ref = model.forward_process(bt["input_ids"].clone(), bt["labels"].clone(), **kw)
    mine = training.forward_process(lambda ids: llada.forward_logits(ids, sd, cfg), bt["input_ids"].clone(),
                                    bt["labels"].clone(), mask_token_id=cfg["mask_token_id"], **kw)
    for a, b, what in zip(ref, mine, ("logits", "loss_t2i", "loss_lm", "loss_mmu")):
        assert torch.equal(a, b), f"restatement != reference (forward_process {what})"
    _save(name, loss_t2i=ref[1], loss_lm=ref[2], loss_mmu=ref[3], logits_sub=ref[0][:, ::5, ::997],
          meta=np.array([B_t2i, B_lm, B_mmu, L, max_seq_length, wseed, seed]))


def forward_process_r2i_case(name, cfg, B_t2i, B_lm, B_mmu, B_r2i, L, max_seq_length, wseed, seed):
    """forward_process_with_r2i (modeling_mmada.py:278-356) and forward_t2i (:359-385): the real methods against
    oracle/training.py.  The batch is training.make_batch's with its last group split into mmu and r2i rows."""
    print(f"[forward_process_with_r2i / forward_t2i] {name}")
    sd = W.make_llada_weights(cfg, wseed)
    model = rb.build_model(cfg, sd)
    bt = training.make_batch(B_t2i, B_lm, B_mmu + B_r2i, L, max_seq_length, seed, cfg["mask_token_id"])
    kw = dict(t2i_masks=bt["t2i_masks"], max_seq_length=max_seq_length, batch_size_t2i=B_t2i, batch_size_lm=B_lm,
              batch_size_mmu=B_mmu, batch_size_r2i=B_r2i, p_mask_lm=bt["p_mask_lm"], p_mask_mmu=bt["p_mask_mmu"][:B_mmu],
              p_mask_r2i=bt["p_mask_mmu"][B_mmu:], answer_lengths=bt["answer_lengths"][:B_mmu],
              answer_lengths_lm=bt["answer_lengths_lm"], answer_lengths_r2i=bt["answer_lengths"][B_mmu:])
    with torch.no_grad():
        ref = model.forward_process_with_r2i(bt["input_ids"].clone(), bt["labels"].clone(), **kw)
        ref_t2i = model.forward_t2i(bt["input_ids"].clone(), bt["labels"].clone(), batch_size_t2i=B_t2i,
                                    max_seq_length=max_seq_length, t2i_masks=bt["t2i_masks"])
    fn = lambda ids: llada.forward_logits(ids, sd, cfg)
    mine = training.forward_process_with_r2i(fn, bt["input_ids"].clone(), bt["labels"].clone(),
                                             mask_token_id=cfg["mask_token_id"], **kw)
    for a, b, what in zip(ref, mine, ("logits", "loss_t2i", "loss_lm", "loss_mmu", "loss_r2i")):
        assert torch.equal(a, b), f"restatement != reference (forward_process_with_r2i {what})"
    mine_t2i = training.forward_t2i(fn, bt["input_ids"].clone(), bt["labels"].clone(), batch_size_t2i=B_t2i,
                                    max_seq_length=max_seq_length, t2i_masks=bt["t2i_masks"])
    assert torch.equal(ref_t2i, mine_t2i), "restatement != reference (forward_t2i)"
    _save(name, loss_t2i=ref[1], loss_lm=ref[2], loss_mmu=ref[3], loss_r2i=ref[4], loss_forward_t2i=ref_t2i,
          meta=np.array([B_t2i, B_lm, B_mmu, B_r2i, L, max_seq_length, wseed, seed]))


def prompting_case():
    """training/prompting_utils.py (the real class, stub tokenizer) against oracle/prompting.py; stores ragged inputs and
    the reference's outputs."""
    import importlib.util
    from . import prompting as OP
    print("[prompting]")
    spec = importlib.util.spec_from_file_location("ref_prompting_utils", os.path.join(rb.REF, "training", "prompting_utils.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)

    class Tok:
        bos_token_id, eos_token_id = OP.BOS, OP.EOS
        _special = {"<|end_header_id|>": OP.END_HEADER, "<|eot_id|>": 126348, "<|start_header_id|>": 126346}

        def __init__(self):
            self.table = {}

        def __len__(self):
            return W.TEXT_VOCAB

        def convert_tokens_to_ids(self, toks):
            return [self._special[t] for t in toks]

        def __call__(self, texts, **kw):
            return {"input_ids": [list(self.table[t]) for t in texts]}

    g = torch.Generator().manual_seed(77)
    out = {}
    for name, mtl, N, lens in (("t2i_a", 16, 8, [0, 1, 5, 13, 14, 15, 16, 30]),      # padding, exact fit, truncation
                                ("t2i_b", 512, 1024, [0, 7, 64, 200])):
        with rb.quiet():
            up = ref.UniversalPrompting(Tok(), max_text_len=mtl, use_reserved_token=True)
        texts = [torch.randint(0, 126000, (n,), generator=g).tolist() for n in lens]
        texts[2] = [OP.BOS] + texts[2][1:] if texts[2] else texts[2]                 # a text that already starts with bos
        image = torch.randint(W.TEXT_VOCAB, W.TEXT_VOCAB + 8192, (len(lens), N), generator=g)
        image[0, :] = 126336
        ids_ref, mask_ref = up.t2i_gen_prompt([list(t) for t in texts], image)
        # through __call__ with the stub tokenizer, as inference_t2i.py:92 does
        up.text_tokenizer.table = {f"p{i}": t for i, t in enumerate(texts)}
        ids_call, mask_call = up(([f"p{i}" for i in range(len(texts))], image), "t2i_gen")
        assert torch.equal(ids_ref, ids_call) and torch.equal(mask_ref, mask_call)
        for i, t in enumerate(texts):
            ids_o, mask_o = OP.prefix_layout(t, image[i].tolist(), mtl, OP.RESERVED["<|t2i|>"], OP.RESERVED["<|soi|>"],
                                             OP.RESERVED["<|eoi|>"])
            assert ids_o == ids_ref[i].tolist() and mask_o == mask_ref[i].tolist(), (name, i)
        # t2m layout (cond_dropout_prob = 0: inference)
        with rb.quiet():
            upm = ref.UniversalPrompting(Tok(), max_text_len=mtl, use_reserved_token=True, cond_dropout_prob=0.0)
        motion = torch.randint(W.TEXT_VOCAB + 8192, W.TEXT_VOCAB + 8192 + 512, (len(lens), N), generator=g)
        ids_m, mask_m, _ = upm.t2m_prompt([list(t) for t in texts], motion, motion.clone())
        for i, t in enumerate(texts):
            ids_o, mask_o = OP.prefix_layout(t, motion[i].tolist(), mtl, OP.RESERVED["<|t2m|>"], OP.RESERVED["<|som|>"],
                                             OP.RESERVED["<|eom|>"])
            assert ids_o == ids_m[i].tolist() and mask_o == mask_m[i].tolist(), (name, "t2m", i)
        flat = torch.tensor([x for t in texts for x in t], dtype=torch.int64)
        off = torch.tensor([0] + list(np.cumsum([len(t) for t in texts])), dtype=torch.int64)
        out.update({f"{name}_text": flat, f"{name}_off": off, f"{name}_image": image, f"{name}_ids": ids_ref,
                    f"{name}_mask": mask_ref, f"{name}_max_text_len": mtl, f"{name}_motion": motion,
                    f"{name}_t2m_ids": ids_m, f"{name}_t2m_mask": mask_m})
    # mmu_gen_prompt: one row per call (its prompt masks only concatenate for equal prompt lengths)
    mtl, N = 24, 16
    with rb.quiet():
        up = ref.UniversalPrompting(Tok(), max_text_len=mtl, use_reserved_token=True)
    lens = [0, 3, 10, 23, 24, 40]
    texts = [torch.randint(0, 126000, (n,), generator=g).tolist() for n in lens]
    texts[2][4] = OP.END_HEADER
    texts[3][1] = OP.END_HEADER; texts[3][9] = OP.END_HEADER
    texts[5][30] = OP.END_HEADER                                                     # beyond the truncation point
    image = torch.randint(W.TEXT_VOCAB, W.TEXT_VOCAB + 8192, (len(lens), N), generator=g)
    seqs, plens = [], []
    for i, t in enumerate(texts):
        with rb.quiet():
            ids_ref, pm_ref = up.mmu_gen_prompt(image[i:i + 1], [list(t)])
        seq_o, plen_o = OP.mmu_gen_layout(t, image[i].tolist(), mtl)
        assert seq_o == ids_ref[0].tolist(), ("mmu_gen", i)
        assert OP.mmu_gen_mask(plen_o, mtl) == pm_ref[0].tolist(), ("mmu_gen mask", i)
        seqs.append(ids_ref[0]); plens.append(plen_o)
    out.update({"mmu_text": torch.tensor([x for t in texts for x in t], dtype=torch.int64),
                "mmu_off": torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int64), "mmu_image": image,
                "mmu_ids": torch.stack(seqs), "mmu_prompt_length": torch.tensor(plens), "mmu_max_text_len": mtl,
                "end_header": OP.END_HEADER})
    _save("prompting", **out)


def main():
    assert rb.available(), "needs /root/reference (build container)"
    torch.set_num_threads(os.cpu_count())
    sampling_case()
    magvit_case()
    magvit_encoder_case()
    motion_case()
    prompting_case()
    forward_process_case("forward_process_tiny", W.TINY, B_t2i=2, B_lm=2, B_mmu=2, L=96, max_seq_length=31, wseed=0, seed=41)
    forward_process_r2i_case("forward_process_r2i_tiny", W.TINY, B_t2i=2, B_lm=1, B_mmu=2, B_r2i=2, L=80, max_seq_length=23,
                             wseed=0, seed=43)
    logits_case("logits_tiny", W.TINY, 2, 96, 0, 5)
    logits_case("logits_tiny128", W.TINY128, 2, 200, 1, 6)
    t2i_case("t2i_tiny", W.TINY, B=2, P=33, N=64, steps=15, guidance=3.5, wseed=0, pseed=1, gseed=1234)
    t2i_case("t2i_tiny128", W.TINY128, B=2, P=17, N=256, steps=8, guidance=2.0, wseed=1, pseed=2, gseed=99)
    text_case("text_t0", W.TINY, B=3, Lp=12, gen=32, block=8, steps=16, temperature=0.0, cfg_scale=0.0,
              wseed=0, seed=21, fast_eot="hit")
    text_case("text_t1", W.TINY, B=3, Lp=12, gen=32, block=8, steps=16, temperature=1.0, cfg_scale=0.0,
              wseed=0, seed=22, fast_eot="miss")
    text_case("text_cfg", W.TINY, B=2, Lp=10, gen=32, block=16, steps=8, temperature=0.7, cfg_scale=1.5,
              wseed=0, seed=23)
    t2m_case("t2m_tiny", W.TINY_T2M, B=2, Lt=20, N=64, steps=6, wseed=0, seed=31, gseed=77)
    if "--skip-c1" not in sys.argv:
        t2i_case("t2i_c1", W.C1, B=1, P=129, N=256, steps=15, guidance=3.5, wseed=0, pseed=0, gseed=1234)


if __name__ == "__main__":
    main()

"""CPU: the oracle restatement reproduces the golden vectors produced by the REAL reference
(oracle/make_goldens.py).  Inputs (weights, prompts, noise) are regenerated from seeds."""
import numpy as np
import torch

from oracle import denoise, llada, magvit, weights as W


def test_sampling_tables_and_topk(golden):
    gd = golden("sampling")
    for N in (64, 256, 1024):
        for T in (8, 15, 18):
            assert denoise.t2i_mask_len_schedule(N, T) == list(gd[f"cos_{N}_{T}"])
    # Appendix C: fp32 host cosine gives 511 (not 512) at s=9 and -1 (not 0) at s=14
    assert gd["cos_1024_15"][9] == 511.0 and gd["cos_1024_15"][14] == -1.0
    probs, u, ml = (torch.from_numpy(gd[k]) for k in ("probs", "u", "mask_len"))
    for i, T in enumerate(gd["temperatures"]):
        got = denoise.mask_by_random_topk(ml, probs, float(T), u=u)
        assert np.array_equal(got.numpy(), gd["masking"][i])
        # generator path draws the same noise
        got2 = denoise.mask_by_random_topk(ml, probs, float(T), generator=torch.Generator().manual_seed(11))
        assert np.array_equal(got2.numpy(), gd["masking"][i])


def test_logits_tiny(golden):
    gd = golden("logits_tiny")
    sd = W.make_llada_weights(W.TINY, int(gd["meta"][2]))
    lg = llada.forward_logits(torch.from_numpy(gd["ids"]), sd, W.TINY)
    cols = slice(W.TEXT_VOCAB, W.TEXT_VOCAB + W.CODEBOOK)
    assert np.array_equal(lg[:, ::3, cols][:, :, ::16].numpy(), gd["logits_img"])
    assert np.array_equal(lg[:, ::3, :126000:512].numpy(), gd["logits_txt"])


def test_t2i_tiny_end_to_end(golden):
    gd = golden("t2i_tiny")
    B, P, N, steps, wseed, pseed, gseed = (int(v) for v in gd["meta"])
    sd = W.make_llada_weights(W.TINY, wseed)
    cond, unc, _, _ = W.make_t2i_prompts(B, P, N, pseed)
    assert np.array_equal(cond.numpy(), gd["cond_ids"]) and np.array_equal(unc.numpy(), gd["uncond_ids"])
    ids = cond.clone()
    trace = []
    out = denoise.t2i_generate(lambda x: llada.forward_logits(x, sd, W.TINY), ids, unc.clone(),
                               guidance_scale=float(gd["guidance"]), timesteps=steps, seq_len=N, resolution=P - 1,
                               generator=torch.Generator().manual_seed(gseed), trace=trace)
    assert np.array_equal(out.numpy(), gd["sampled_ids"])
    assert np.array_equal(ids.numpy(), gd["final_input_ids"])
    assert np.array_equal(torch.stack([t["masking"] for t in trace]).numpy(), gd["step_masking"])
    assert np.array_equal(torch.stack([t["sampled_ids"] for t in trace]).numpy(), gd["step_sampled"])
    # Q9: exactly one token stays masked in the caller's ids, the returned ids are complete
    assert int((ids == 126336).sum()) == B and int((out == 126336).sum()) == 0
    # Q3: compounding temperature
    t = 1.0
    for s in range(steps):
        t *= 1.0 - (s + 1) / steps
        assert abs(t - float(gd["step_temperature"][s])) < 1e-12


def test_text_generation(golden):
    for name in ("text_t0", "text_t1", "text_cfg"):
        gd = golden(name)
        B, Lp, gen, block, steps, wseed, seed = (int(v) for v in gd["meta"])
        sd = W.make_llada_weights(W.TINY, wseed)
        prompt = torch.from_numpy(gd["prompt"])
        torch.manual_seed(seed)
        x = denoise.generate(lambda t: llada.forward_logits(t, sd, W.TINY), prompt, steps=steps, gen_length=gen,
                             block_length=block, temperature=float(gd["temperature"]), cfg_scale=float(gd["cfg_scale"]))
        assert np.array_equal(x.numpy(), gd["x"]), name
        if "fast_x" in gd.files:
            torch.manual_seed(seed)
            xf = denoise.generate(lambda t: llada.forward_logits(t, sd, W.TINY), prompt[:1], steps=steps, gen_length=gen,
                                  block_length=block, temperature=float(gd["temperature"]),
                                  cfg_scale=float(gd["cfg_scale"]), eot_token=int(gd["fast_eot"]))
            assert np.array_equal(xf.numpy(), gd["fast_x"]), name


def test_get_num_transfer_tokens():
    m = torch.zeros(3, 16, dtype=torch.bool)
    m[0, :7] = True; m[1, :] = True
    out = denoise.get_num_transfer_tokens(m, 4)
    assert out.tolist() == [[2, 2, 2, 1], [4, 4, 4, 4], [0, 0, 0, 0]]


def test_magvit(golden):
    gd = golden("magvit")
    bits = magvit.lfq_indices_to_bits(gd["bits_sample_idx"][None])
    assert np.array_equal(bits.reshape(13, 4).T, gd["bits_sample"].reshape(13, 4).T)
    # 5 = ...0101b -> last three channels +1 -1 +1 (MSB first)
    assert list(bits[0, -3:, 0, 1]) == [1.0, -1.0, 1.0]
    allidx = np.arange(8192).reshape(8, 1024)
    assert np.array_equal(magvit.lfq_bits_to_indices(magvit.lfq_indices_to_bits(allidx)).reshape(8, 1024), allidx)
    sd = W.make_vq_decoder_weights(0)
    taps = {}
    pix = magvit.decode_code(torch.from_numpy(gd["idx_8x8"]), sd, taps)
    assert np.array_equal(pix.numpy(), gd["pix_8x8"])
    assert list(taps.keys()) == list(gd["tap_names_8x8"])


def test_t2m_tiny(golden):
    gd = golden("t2m_tiny")
    B, Lt, N, steps, wseed, seed, gseed = (int(v) for v in gd["meta"])
    sd = W.make_llada_weights(W.TINY_T2M, wseed)
    ids = torch.from_numpy(gd["ids"]).clone()
    out = denoise.t2m_generate(lambda x: llada.forward_logits(x, sd, W.TINY_T2M), ids, timesteps=steps, seq_len=N,
                               generator=torch.Generator().manual_seed(gseed), som_token=126096, eom_token=126097)
    assert np.array_equal(out.numpy(), gd["sampled_ids"])
    assert np.array_equal(ids.numpy(), gd["final_input_ids"])
    assert int((ids == 126336).sum()) == 0            # Q15: no re-masking on the last step


def test_magvit_encoder(golden):
    """get_code: the restated VQGAN encoder reproduces the real reference's latents and code ids bit for bit."""
    gd = golden("magvit_encoder")
    sd = W.make_vq_encoder_weights(0)
    px = torch.from_numpy(gd["pixels"]).float()
    with torch.no_grad():
        z = magvit.encoder_forward(px, sd)
    assert np.array_equal(z.numpy(), gd["latents"])
    assert np.array_equal(magvit.lfq_bits_to_indices(z.numpy()).reshape(1, -1), gd["codes"])


def test_motion_decoder(golden):
    """Motion VQ-VAE decode: the restatement reproduces the output of the reference's own Decoder class."""
    from oracle import motion
    gd = golden("motion_decoder")
    sd = motion.make_motion_decoder_weights(0)
    with torch.no_grad():
        pose = motion.forward_decoder(torch.from_numpy(gd["ids"]), sd)
    assert pose.shape == (1, 49 * 4, 263)
    assert np.array_equal(pose.numpy(), gd["pose"])


def test_forward_process_oracle_reproduces_reference_losses(golden):
    """oracle/training.py on the synthetic mixed batch: the losses stored from the REAL reference's forward_process,
    bit for bit (same torch ops on the same fp32 logits)."""
    from oracle import llada, training, weights as W
    gd = golden("forward_process_tiny")
    B_t2i, B_lm, B_mmu, L, msl, wseed, seed = (int(v) for v in gd["meta"])
    sd = W.make_llada_weights(W.TINY, wseed)
    bt = training.make_batch(B_t2i, B_lm, B_mmu, L, msl, seed, W.TINY["mask_token_id"])
    lg, l_t2i, l_lm, l_mmu = training.forward_process(
        lambda ids: llada.forward_logits(ids, sd, W.TINY), bt["input_ids"], bt["labels"], batch_size_t2i=B_t2i,
        batch_size_lm=B_lm, batch_size_mmu=B_mmu, max_seq_length=msl, p_mask_lm=bt["p_mask_lm"], p_mask_mmu=bt["p_mask_mmu"],
        answer_lengths=bt["answer_lengths"], t2i_masks=bt["t2i_masks"], answer_lengths_lm=bt["answer_lengths_lm"],
        mask_token_id=W.TINY["mask_token_id"])
    assert np.array_equal(lg[:, ::5, ::997].numpy(), gd["logits_sub"])
    assert float(l_t2i) == float(gd["loss_t2i"]) and float(l_lm) == float(gd["loss_lm"]) and float(l_mmu) == float(gd["loss_mmu"])


def test_forward_process_r2i_oracle_reproduces_reference_losses(golden):
    """oracle/training.py: forward_process_with_r2i and forward_t2i against the REAL methods' stored losses, bit for bit."""
    from oracle import llada, training, weights as W
    gd = golden("forward_process_r2i_tiny")
    B_t2i, B_lm, B_mmu, B_r2i, L, msl, wseed, seed = (int(v) for v in gd["meta"])
    sd = W.make_llada_weights(W.TINY, wseed)
    bt = training.make_batch(B_t2i, B_lm, B_mmu + B_r2i, L, msl, seed, W.TINY["mask_token_id"])
    fn = lambda ids: llada.forward_logits(ids, sd, W.TINY)
    out = training.forward_process_with_r2i(
        fn, bt["input_ids"], bt["labels"], t2i_masks=bt["t2i_masks"], max_seq_length=msl, batch_size_t2i=B_t2i,
        batch_size_lm=B_lm, batch_size_mmu=B_mmu, batch_size_r2i=B_r2i, p_mask_lm=bt["p_mask_lm"],
        p_mask_mmu=bt["p_mask_mmu"][:B_mmu], p_mask_r2i=bt["p_mask_mmu"][B_mmu:], answer_lengths=bt["answer_lengths"][:B_mmu],
        answer_lengths_lm=bt["answer_lengths_lm"], answer_lengths_r2i=bt["answer_lengths"][B_mmu:],
        mask_token_id=W.TINY["mask_token_id"])
    for v, name in zip(out[1:], ("loss_t2i", "loss_lm", "loss_mmu", "loss_r2i")):
        assert float(v) == float(gd[name]), name
    t2i = training.forward_t2i(fn, bt["input_ids"], bt["labels"], batch_size_t2i=B_t2i, max_seq_length=msl, t2i_masks=bt["t2i_masks"])
    assert float(t2i) == float(gd["loss_forward_t2i"])

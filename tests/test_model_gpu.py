"""Model-level parity on the GPU: the LLaDA forward against the reference's fp32 logits (golden
fixtures produced by the real reference, oracle/make_goldens.py) and the t2i loop's decisions
against the oracle on identical logits and noise."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 2e-2      # north_star: max |delta| / max |ref| <= 2e-2 against the reference's fp32 path


class _UP:
    class _T:
        def __len__(self):
            return 126349
    text_tokenizer = _T()


def _model(cfg_dict, seed, fused_norm=None):
    from mmada_b200 import MMadaConfig, MMadaModelLM
    from oracle import weights as W
    m = MMadaModelLM(MMadaConfig.from_dict(cfg_dict), fused_norm=fused_norm)
    m.load_state_dict(W.make_llada_weights(cfg_dict, seed))
    return m


@pytest.mark.parametrize("fused_norm", [True, False], ids=["norm_folded", "norm_kernel"])
@pytest.mark.parametrize("name,cfgname,wseed", [("logits_tiny", "TINY", 0), ("logits_tiny128", "TINY128", 1)])
def test_forward_logits_vs_reference_fp32(golden, name, cfgname, wseed, fused_norm):
    """Both block pipelines — RMSNorm folded into the GEMMs around it (the default) and the stand-alone RMSNorm
    kernel — against the reference's fp32 logits."""
    from oracle import weights as W
    gd = golden(name)
    m = _model(getattr(W, cfgname), wseed, fused_norm)
    assert m.fused_norm == fused_norm
    ids = torch.from_numpy(gd["ids"]).cuda()
    lg = m(ids).logits.float().cpu()
    cols = slice(W.TEXT_VOCAB, W.TEXT_VOCAB + W.CODEBOOK)
    mine_img = lg[:, ::3, cols][:, :, ::16]
    mine_txt = lg[:, ::3, :126000:512]
    scale = float(gd["absmax"])
    err = max(float((mine_img - torch.from_numpy(gd["logits_img"])).abs().max()),
              float((mine_txt - torch.from_numpy(gd["logits_txt"])).abs().max())) / scale
    print(f"{name}: max|d|/max|ref| = {err:.3e}")
    assert err < TOL


@pytest.mark.parametrize("fused_norm", [True, False], ids=["norm_folded", "norm_kernel"])
def test_last_block_on_requested_rows_is_bit_identical(fused_norm):
    """logits_rows(rows=...) gathers the requested token rows after the last block's attention and runs attn_out, the
    MLP, ln_f and the head on them only: every value must equal the all-rows computation bit for bit."""
    from oracle import weights as W
    m = _model(W.TINY128, 1, fused_norm)
    g = torch.Generator().manual_seed(3)
    B, L = 3, 200
    ids = torch.randint(0, 126000, (B, L), generator=g).cuda()
    rows = torch.cat([torch.arange(b * L + 70, b * L + 170) for b in range(B)]).to(torch.int32).cuda()     # ragged vs tiles
    n0 = m.kernel_launches
    some = m.logits_rows(ids, rows, 100, 100 + 1024)
    n_some = m.kernel_launches - n0
    m.restrict_last_block = False
    same = m.logits_rows(ids, rows, 100, 100 + 1024)
    full = m.logits_rows(ids, None, 100, 100 + 1024)
    assert n_some == (m.kernel_launches - n0 - n_some) // 2 + 2            # the two gathers
    assert torch.equal(some, same) and torch.equal(some, full[rows.long()])


@pytest.mark.parametrize("name,cfgname", [("t2i_tiny", "TINY"), ("t2i_tiny128", "TINY128"), ("t2i_c1", "C1")])
def test_t2i_generate_decisions_match_oracle(golden, name, cfgname):
    """Feed the CUDA path's own fp32 logits and the same noise to the CPU oracle step by step: sampled
    ids, masks and the carried state must be bit-identical.  Also reports agreement with the
    reference's end-to-end golden run (bf16 weights vs fp32, so not asserted to be exact)."""
    from oracle import denoise, weights as W
    gd = golden(name)
    B, P, N, steps, wseed, pseed, gseed = (int(v) for v in gd["meta"])
    guidance = float(gd["guidance"])
    m = _model(getattr(W, cfgname), wseed)
    cond, unc, _, _ = W.make_t2i_prompts(B, P, N, pseed)
    assert np.array_equal(cond.numpy(), gd["cond_ids"])
    # the reference's noise stream, regenerated from the golden run's CPU generator seed
    g = torch.Generator().manual_seed(gseed)
    noise = []
    for s in range(steps):
        q = torch.empty(B * N, W.CODEBOOK).exponential_(1, generator=g)
        u = torch.zeros(B, N).uniform_(0, 1, generator=g)
        assert abs(float(q.double().sum()) - float(gd["step_q_sum"][s])) < 1e-6 * abs(float(gd["step_q_sum"][s]))
        assert np.array_equal(u.numpy(), gd["step_u"][s])
        noise.append((q, u))
    ids = cond.clone().cuda()
    trace = []
    out = m.t2i_generate(input_ids=ids, uncond_input_ids=unc.cuda(), guidance_scale=guidance, timesteps=steps, seq_len=N,
                         resolution=P - 1, uni_prompting=_UP(), noise=noise, trace=trace)
    known = torch.full((B, N), 126336, dtype=torch.int64)
    temperature = 1.0
    for s, t in enumerate(trace):
        temperature = temperature * (1.0 - (s + 1) / steps)
        r = denoise.t2i_sample_step(t["cond"].cpu(), t["uncond"].cpu(), guidance, known, 126336,
                                    float(gd["step_mask_len_raw"][s]), temperature, *noise[s])
        assert torch.equal(t["sampled_ids"].cpu(), r["sampled_ids"]), f"step {s}: sampled ids differ"
        assert torch.equal(t["masking"].cpu(), r["masking"]), f"step {s}: masking differs"
        known = r["next_known"]
    final = torch.where(r["masking"], 126336, r["sampled_ids"] + W.TEXT_VOCAB)
    assert torch.equal(ids[:, -(N + 1):-1].cpu(), final)
    assert torch.equal(out.cpu(), r["sampled_ids"])
    # logits of the first forward against the reference's fp32 logits
    ref0 = torch.from_numpy(gd["first_logits_sub"])
    mine0 = torch.cat([trace[0]["cond"], trace[0]["uncond"]]).cpu()[:, ::4, ::32]
    err = float((mine0 - ref0).abs().max()) / float(gd["first_logits_absmax"])
    agree0 = float((trace[0]["sampled_ids"].cpu() == torch.from_numpy(gd["step_sampled"][0])).float().mean())
    agree = float((out.cpu() == torch.from_numpy(gd["sampled_ids"])).float().mean())
    print(f"{name}: first-step logits err {err:.3e}; sampled-id agreement with the fp32 reference run: "
          f"step0 {agree0:.3f}, final {agree:.3f}")
    assert err < TOL


def test_t2m_generate_decisions_match_oracle(golden):
    """t2m_generate (reference fork models/modelling_ours.py:557-682): raw last-step samples returned,
    no re-masking on the last step, non-compounding temperature — replayed on the oracle with the CUDA
    path's logits and the same noise."""
    from oracle import denoise, weights as W
    gd = golden("t2m_tiny")
    B, Lt, N, steps, wseed, seed, gseed = (int(v) for v in gd["meta"])
    m = _model(W.TINY_T2M, wseed)
    ids0 = torch.from_numpy(gd["ids"])
    g = torch.Generator().manual_seed(gseed)
    noise = []
    for s in range(steps):
        q = torch.empty(B * N, 512).exponential_(1, generator=g)
        u = torch.zeros(B, N).uniform_(0, 1, generator=g) if s < steps - 1 else torch.zeros(B, N)
        noise.append((q, u))

    class UP(_UP):
        sptids_dict = {"<|som|>": torch.tensor([126096]), "<|eom|>": torch.tensor([126097])}

    ids = ids0.clone().cuda()
    out = m.t2m_generate(input_ids=ids, timesteps=steps, seq_len=N, uni_prompting=UP(), noise=noise)
    # oracle replay: logits from the CUDA model itself (full forward, sliced), same noise
    off = W.TEXT_VOCAB + 8192
    ido = ids0.clone()
    outo = denoise.t2m_generate(lambda x: m(x.cuda()).logits.float().cpu(), ido, timesteps=steps, seq_len=N,
                                som_token=126096, eom_token=126097, noise=noise)
    assert torch.equal(out.cpu(), outo)
    assert torch.equal(ids.cpu(), ido)
    assert int((ids == 126336).sum()) == 0
    assert out.min() >= 0 and out.max() < 512
    agree = float((out.cpu() == torch.from_numpy(gd["sampled_ids"])).float().mean())
    print(f"t2m_tiny: agreement with the fp32 reference run {agree:.3f}")


def test_t2i_stepwise_yields_images():
    from mmada_b200.modeling_magvitv2 import MAGVITv2
    from oracle import weights as W
    m = _model(W.TINY, 0)
    vq = MAGVITv2().load_state_dict(W.make_vq_decoder_weights(0))
    B, P, N, steps = 1, 17, 256, 3
    cond, unc, _, _ = W.make_t2i_prompts(B, P, N, 3)
    g1 = torch.Generator(device="cuda").manual_seed(5)
    ids1 = cond.clone().cuda()
    frames = list(m.t2i_generate_decoding_stepwise(input_ids=ids1, uncond_input_ids=unc.cuda(), guidance_scale=2.0,
                                                   timesteps=steps, seq_len=N, resolution=P - 1, generator=g1,
                                                   uni_prompting=_UP(), vq_model=vq))
    assert [t for _, t in frames] == [f"Step {i + 1}/{steps}" for i in range(steps)]
    assert frames[0][0].size == (256, 256) and frames[0][0].mode == "RGB"
    # same generator seed through t2i_generate: identical final state
    ids2 = cond.clone().cuda()
    out = m.t2i_generate(input_ids=ids2, uncond_input_ids=unc.cuda(), guidance_scale=2.0, timesteps=steps, seq_len=N,
                         resolution=P - 1, generator=torch.Generator(device="cuda").manual_seed(5), uni_prompting=_UP())
    assert torch.equal(ids1, ids2)
    import numpy as np
    last = np.asarray(frames[-1][0])
    assert np.array_equal(last, vq.decode_code_uint8(out[:1].clamp(0, 8191))[0].cpu().numpy())


def test_cross_entropy_rows_vs_torch():
    """mmada_cross_entropy_rows_f32 against F.cross_entropy(reduction='none') on the same fp32 logits (tolerance 2e-6
    relative: order of the fp32 sum of exponentials), ignore_index rows, a non-multiple-of-4 vocabulary, strided rows."""
    import torch.nn.functional as F
    from mmada_b200 import ops
    g = torch.Generator().manual_seed(3)
    for R, V in ((37, 134656), (5, 1003), (64, 8192)):
        lg = (torch.randn(R, V, generator=g) * 4).cuda()
        lab = torch.randint(0, V, (R,), generator=g).cuda()
        lab[::7] = -100
        ref = F.cross_entropy(lg.double(), lab, ignore_index=-100, reduction="none").float()
        out = ops.cross_entropy_rows(lg, lab, -100)
        assert torch.equal(out[::7], torch.zeros_like(out[::7]))
        torch.testing.assert_close(out, ref, rtol=2e-6, atol=2e-6)
    wide = (torch.randn(9, 4096 + 64, generator=g)).cuda()
    lab = torch.randint(0, 4096, (9,), generator=g).cuda()
    torch.testing.assert_close(ops.cross_entropy_rows(wide[:, :4096], lab), F.cross_entropy(wide[:, :4096], lab, reduction="none"),
                               rtol=2e-6, atol=2e-6)


def test_forward_process_matches_reference_golden(golden):
    """MMadaModelLM.forward_process (modeling_mmada.py:213-276, forward values) on a mixed t2i / lm / mmu batch:
    (a) the loss reductions replayed by the CPU oracle on the CUDA path's own full logits — 1e-5 relative (fp32
    summation order); (b) against the REAL reference's fp32 losses from tests/golden — 2e-2 relative (bf16 weights and
    activations; north_star's logit tolerance)."""
    from mmada_b200 import MMadaConfig, MMadaModelLM
    from oracle import training, weights as W
    gd = golden("forward_process_tiny")
    B_t2i, B_lm, B_mmu, L, msl, wseed, seed = (int(v) for v in gd["meta"])
    m = MMadaModelLM(MMadaConfig.from_dict(W.TINY)).load_state_dict(W.make_llada_weights(W.TINY, wseed))
    bt = training.make_batch(B_t2i, B_lm, B_mmu, L, msl, seed, W.TINY["mask_token_id"])
    kw = dict(batch_size_t2i=B_t2i, batch_size_lm=B_lm, batch_size_mmu=B_mmu, max_seq_length=msl, p_mask_lm=bt["p_mask_lm"],
              p_mask_mmu=bt["p_mask_mmu"], answer_lengths=bt["answer_lengths"], t2i_masks=bt["t2i_masks"],
              answer_lengths_lm=bt["answer_lengths_lm"])
    logits, l_t2i, l_lm, l_mmu = m.forward_process(bt["input_ids"].cuda(), bt["labels"].cuda(), return_logits=True, **kw)
    assert logits.shape == (B_t2i + B_lm + B_mmu, L, W.TINY["vocab_size"])
    none_logits = m.forward_process(bt["input_ids"].cuda(), bt["labels"].cuda(), **kw)
    assert none_logits[0] is None and all(torch.equal(a, b) for a, b in zip(none_logits[1:], (l_t2i, l_lm, l_mmu)))
    full = logits.cpu()
    _, o_t2i, o_lm, o_mmu = training.forward_process(lambda ids: full, bt["input_ids"], bt["labels"],
                                                     mask_token_id=W.TINY["mask_token_id"], **kw)
    for mine, orc, ref, name in ((l_t2i, o_t2i, gd["loss_t2i"], "t2i"), (l_lm, o_lm, gd["loss_lm"], "lm"),
                                 (l_mmu, o_mmu, gd["loss_mmu"], "mmu")):
        a, b, c = float(mine), float(orc), float(ref)
        print(f"forward_process loss_{name}: cuda {a:.6f}  oracle on the same logits {b:.6f}  reference fp32 {c:.6f}")
        assert abs(a - b) <= 1e-5 * abs(b)
        assert abs(a - c) <= 2e-2 * abs(c)
    sub = torch.from_numpy(gd["logits_sub"])
    err = float((full[:, ::5, ::997] - sub).abs().max() / sub.abs().max())
    assert err <= 2e-2
    # batch_size_mmu = 0: the reference's [-0:] slices take the whole batch (kept)
    z = dict(kw, batch_size_mmu=0, p_mask_mmu=torch.full((B_t2i + B_lm + B_mmu, L), 0.5),
             answer_lengths=torch.full((B_t2i + B_lm + B_mmu, L), 7))
    _, _, _, q_mmu = m.forward_process(bt["input_ids"].cuda(), bt["labels"].cuda(), **z)
    _, _, _, qo_mmu = training.forward_process(lambda ids: full, bt["input_ids"], bt["labels"],
                                               mask_token_id=W.TINY["mask_token_id"], **z)
    assert abs(float(q_mmu) - float(qo_mmu)) <= 1e-5 * abs(float(qo_mmu))


def test_logits_rows_cuda_graph_replay_is_bit_identical():
    """Small workloads replay logits_rows from a CUDA graph (one per shape): same bits as the eager launches, for new
    ids / row lists of the same shape, after a flag change (its own graph) and after reloading weights (graphs dropped)."""
    from mmada_b200 import MMadaConfig, MMadaModelLM
    from oracle import weights as W
    m = MMadaModelLM(MMadaConfig.from_dict(W.TINY)).load_state_dict(W.make_llada_weights(W.TINY, 0))
    g = torch.Generator().manual_seed(0)
    B, L = 2, 70
    outs = []
    for rep in range(3):
        ids = torch.randint(0, 126000, (B, L), generator=g).cuda()
        rows = torch.randperm(B * L, generator=g)[:40].sort().values.to(torch.int32).cuda()
        m.graph_max_token_rows = 0
        eager = m.logits_rows(ids, rows, 100, 100 + 512)
        m.graph_max_token_rows = 8192
        n0 = m.kernel_launches
        graphed = m.logits_rows(ids, rows, 100, 100 + 512)
        assert torch.equal(eager, graphed) and m.kernel_launches > n0
        outs.append(graphed)
    assert len(m._graphs) == 1 and not torch.equal(outs[0], outs[1])             # one shape, one graph; results are copies
    m.restrict_last_block = False
    assert torch.equal(m.logits_rows(ids, rows, 100, 100 + 512), eager) and len(m._graphs) == 2
    m.load_state_dict(W.make_llada_weights(W.TINY, 1))
    assert len(m._graphs) == 0
    m.graph_max_token_rows = 0
    other = m.logits_rows(ids, rows, 100, 100 + 512)
    m.graph_max_token_rows = 8192
    assert torch.equal(m.logits_rows(ids, rows, 100, 100 + 512), other) and not torch.equal(other, eager)


@pytest.mark.parametrize("B,frac_known", [(1, 0.0), (3, 0.4)])
def test_t2i_generate_without_cfg_and_with_known_tokens(B, frac_known):
    """t2i_generate with guidance_scale = 0 (no uncond branch, modeling_mmada.py:158-171 else-path), batch 1, and with part
    of the image tokens already known at entry (the loop keeps them, :183-184): every step replayed on the oracle."""
    from oracle import denoise, weights as W
    cfg = W.TINY128
    m = _model(cfg, 1)
    P, N, steps = 9, 160, 5
    cond, _, _, _ = W.make_t2i_prompts(B, P, N, 3)
    g = torch.Generator().manual_seed(11)
    known0 = torch.full((B, N), 126336, dtype=torch.int64)
    kn = torch.rand(B, N, generator=g) < frac_known
    known0[kn] = torch.randint(0, W.CODEBOOK, (int(kn.sum()),), generator=g)
    cond[:, -(N + 1):-1] = torch.where(known0 == 126336, 126336, known0 + W.TEXT_VOCAB)
    noise = [(torch.empty(B * N, W.CODEBOOK).exponential_(1, generator=g), torch.rand(B, N, generator=g)) for _ in range(steps)]
    ids = cond.clone().cuda()
    trace = []
    out = m.t2i_generate(input_ids=ids, uncond_input_ids=None, guidance_scale=0.0, timesteps=steps, seq_len=N,
                         resolution=P - 1, uni_prompting=_UP(), noise=noise, trace=trace)
    known, temperature = known0.clone(), 1.0
    sched = denoise.t2i_mask_len_schedule(N, steps)
    for s, t in enumerate(trace):
        temperature *= 1.0 - (s + 1) / steps
        assert t["uncond"] is None
        r = denoise.t2i_sample_step(t["cond"].cpu(), None, 0.0, known, 126336, sched[s], temperature, *noise[s])
        assert torch.equal(t["sampled_ids"].cpu(), r["sampled_ids"]) and torch.equal(t["masking"].cpu(), r["masking"]), s
        # tokens known at entry are never resampled
        assert torch.equal(r["sampled_ids"][kn], known0[kn])
        known = r["next_known"]
    assert torch.equal(out.cpu(), r["sampled_ids"])
    assert torch.equal(ids[:, :-(N + 1)].cpu(), cond[:, :-(N + 1)])                    # the prompt is untouched


def test_forward_process_with_r2i_and_forward_t2i_match_reference_golden(golden):
    """forward_process_with_r2i (modeling_mmada.py:278-356) and forward_t2i (:359-385), forward values: against the CPU oracle
    on the CUDA path's own logits (1e-5) and against the REAL reference's fp32 losses (2e-2)."""
    from mmada_b200 import MMadaConfig, MMadaModelLM
    from oracle import training, weights as W
    gd = golden("forward_process_r2i_tiny")
    B_t2i, B_lm, B_mmu, B_r2i, L, msl, wseed, seed = (int(v) for v in gd["meta"])
    m = MMadaModelLM(MMadaConfig.from_dict(W.TINY)).load_state_dict(W.make_llada_weights(W.TINY, wseed))
    bt = training.make_batch(B_t2i, B_lm, B_mmu + B_r2i, L, msl, seed, W.TINY["mask_token_id"])
    kw = dict(t2i_masks=bt["t2i_masks"], max_seq_length=msl, batch_size_t2i=B_t2i, batch_size_lm=B_lm, batch_size_mmu=B_mmu,
              batch_size_r2i=B_r2i, p_mask_lm=bt["p_mask_lm"], p_mask_mmu=bt["p_mask_mmu"][:B_mmu],
              p_mask_r2i=bt["p_mask_mmu"][B_mmu:], answer_lengths=bt["answer_lengths"][:B_mmu],
              answer_lengths_lm=bt["answer_lengths_lm"], answer_lengths_r2i=bt["answer_lengths"][B_mmu:])
    ids, lab = bt["input_ids"].cuda(), bt["labels"].cuda()
    out = m.forward_process_with_r2i(ids, lab, return_logits=True, **kw)
    full = out[0].cpu()
    orc = training.forward_process_with_r2i(lambda x: full, bt["input_ids"], bt["labels"], mask_token_id=W.TINY["mask_token_id"], **kw)
    for mine, o, name in zip(out[1:], orc[1:], ("loss_t2i", "loss_lm", "loss_mmu", "loss_r2i")):
        a, b, c = float(mine), float(o), float(gd[name])
        assert abs(a - b) <= 1e-5 * abs(b) and abs(a - c) <= 2e-2 * abs(c), (name, a, b, c)
    t2i = float(m.forward_t2i(ids, lab, batch_size_t2i=B_t2i, max_seq_length=msl, t2i_masks=bt["t2i_masks"]))
    assert t2i == float(out[1]) and abs(t2i - float(gd["loss_forward_t2i"])) <= 2e-2 * abs(float(gd["loss_forward_t2i"]))

"""Text / MMU path on the GPU: the fp64 Gumbel-max + softmax-confidence kernel and the per-row top-k
transfer against the CPU oracle on identical logits and uniforms; generate()/mmu_generate decisions
step by step; agreement with the reference's golden runs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T,cfg", [(0.0, 0.0), (1.0, 0.0), (0.7, 1.5), (0.0, 2.0)])
def test_text_sample_rows_vs_oracle(T, cfg):
    from mmada_b200 import ops
    from oracle import denoise
    R, V = 37, 134656
    g = torch.Generator().manual_seed(int(T * 10 + cfg * 100))
    l = torch.randn(R, V, generator=g) * 3
    un = torch.randn(R, V, generator=g) * 3 if cfg > 0 else None
    u = torch.rand(R, V, generator=g, dtype=torch.float64) if T > 0 else None
    mixed = un + (cfg + 1) * (l - un) if cfg > 0 else l
    x0_ref, p_ref = denoise.text_sample_rows(mixed, T, u)
    x0, conf = ops.text_sample_rows(l.cuda(), None if un is None else un.cuda(), cfg, T, None if u is None else u.cuda())
    assert torch.equal(x0.cpu(), x0_ref)
    # fp64 softmax probability: CPU vs GPU exp() and summation order -> ~1e-15 relative
    torch.testing.assert_close(conf.cpu(), p_ref, rtol=1e-12, atol=0)


def test_text_sample_inkernel_rng_is_uniform():
    from mmada_b200 import ops
    R, V = 64, 4096
    l = torch.zeros(R, V, device="cuda")
    x0a, _ = ops.text_sample_rows(l, None, 0.0, 1.0, None, seed=1)
    x0b, _ = ops.text_sample_rows(l, None, 0.0, 1.0, None, seed=1)
    x0c, _ = ops.text_sample_rows(l, None, 0.0, 1.0, None, seed=2)
    assert torch.equal(x0a, x0b) and not torch.equal(x0a, x0c)
    assert x0a.min() >= 0 and x0a.max() < V and x0a.unique().numel() > R // 2     # flat logits -> spread-out tokens


def test_transfer_topk_and_counts():
    from mmada_b200 import ops
    B, L, lo, block, steps = 3, 40, 8, 16, 4
    g = torch.Generator().manual_seed(0)
    x = torch.randint(0, 1000, (B, L), generator=g)
    x[0, lo:lo + block] = 126336
    x[1, lo:lo + 7] = 126336
    x[2, lo + 3] = 126336
    xd = x.cuda()
    cnt = ops.block_mask_count(xd, lo, block, 126336)
    assert cnt.tolist() == [16, 7, 1]
    conf = torch.rand(B, block, generator=g, dtype=torch.float64)
    conf[0, 5] = conf[0, 9]                                  # a tie: lower position wins
    x0 = torch.randint(0, 1000, (B, block), generator=g)
    for step in range(steps):
        before = xd.clone().cpu()
        tr = ops.text_transfer(xd, lo, block, x0.cuda(), conf.cuda(), cnt, steps, step, 126336, want_transfer=True).cpu()
        for b in range(B):
            n = int(cnt[b])
            k = n // steps + (1 if step < n % steps else 0)
            masked = before[b, lo:lo + block] == 126336
            c = torch.where(masked, conf[b], torch.tensor(-np.inf, dtype=torch.float64))
            order = sorted(range(block), key=lambda p: (-float(c[p]), p))
            exp = torch.zeros(block, dtype=torch.bool)
            for p in order[:k]:
                if c[p] > -np.inf:
                    exp[p] = True
            assert torch.equal(tr[b], exp), (step, b)
            assert torch.equal(xd[b, lo:lo + block].cpu(), torch.where(exp, x0[b], before[b, lo:lo + block]))
    assert int((xd[:, lo:lo + block] == 126336).sum()) == 0  # everything transferred after `steps` steps


def _model(cfg_dict, seed):
    from mmada_b200 import MMadaConfig, MMadaModelLM
    from oracle import weights as W
    return MMadaModelLM(MMadaConfig.from_dict(cfg_dict)).load_state_dict(W.make_llada_weights(cfg_dict, seed))


@pytest.mark.parametrize("name", ["text_t0", "text_t1", "text_cfg"])
def test_generate_decisions_match_oracle(golden, name):
    import mmada_b200
    from oracle import denoise, weights as W
    gd = golden(name)
    B, Lp, gen, block, steps, wseed, seed = (int(v) for v in gd["meta"])
    T, cfg = float(gd["temperature"]), float(gd["cfg_scale"])
    m = _model(W.TINY, wseed)
    prompt = torch.from_numpy(gd["prompt"])
    V = W.TINY["vocab_size"]
    g = torch.Generator().manual_seed(seed + 1000)
    noise = [torch.rand(B, block, V, generator=g, dtype=torch.float64) for _ in range(steps)] if T > 0 else None
    trace = []
    x = mmada_b200.generate(m, prompt.cuda(), steps=steps, gen_length=gen, block_length=block, temperature=T,
                            cfg_scale=cfg, noise=noise, trace=trace)
    # replay every step on the CPU oracle with the CUDA path's own logits and the same uniforms
    xo = torch.full((B, Lp + gen), 126336, dtype=torch.long)
    xo[:, :Lp] = prompt
    spb = steps // (gen // block)
    for t in trace:
        k, nb, i = t["k"], t["block"], t["step"]
        lo = Lp + nb * block
        if i == 0:
            ntt = denoise.get_num_transfer_tokens(xo[:, lo:lo + block] == 126336, spb)
        lg = t["logits"].cpu()
        n = B * block
        mixed = lg[n:] + (cfg + 1) * (lg[:n] - lg[n:]) if cfg > 0 else lg[:n]
        u = noise[k].reshape(n, V) if noise is not None else None
        x0, p = denoise.text_sample_rows(mixed, T, u)
        assert torch.equal(t["x0"].cpu().reshape(-1), x0), f"forward {k}: tokens differ"
        masked = xo[:, lo:lo + block] == 126336
        conf = torch.where(masked, p.view(B, block), torch.tensor(-np.inf, dtype=torch.float64))
        for j in range(B):
            _, sel = torch.topk(conf[j], k=int(ntt[j, i]))
            xo[j, lo + sel] = x0.view(B, block)[j, sel]
        assert torch.equal(t["x"].cpu(), xo), f"forward {k}: state differs"
    assert torch.equal(x.cpu(), xo)
    assert int((x == 126336).sum()) == 0
    agree = float((x.cpu() == torch.from_numpy(gd["x"])).float().mean())
    print(f"{name}: token agreement with the fp32 reference run: {agree:.3f}")
    # not asserted: with random-init weights the top logits are nearly tied, so bf16-vs-fp32 logit error
    # flips argmaxes and the sequential loop amplifies it; parity is the step-wise check above


def test_mmu_generate_equals_generate_and_fast_exit(golden):
    import mmada_b200
    from oracle import weights as W
    gd = golden("text_t0")
    B, Lp, gen, block, steps, wseed, seed = (int(v) for v in gd["meta"])
    m = _model(W.TINY, wseed)
    prompt = torch.from_numpy(gd["prompt"]).cuda()
    a = mmada_b200.generate(m, prompt, steps=steps, gen_length=gen, block_length=block)
    b = m.mmu_generate(idx=prompt, max_new_tokens=gen, steps=steps, block_length=block)
    assert torch.equal(a, b)
    eot = int(a[0, Lp + block - 1])
    f = m.mmu_generate_fast(idx=prompt[:1], max_new_tokens=gen, steps=steps, block_length=block, eot_token=eot)
    assert torch.equal(f[:, :Lp + block], a[:1, :Lp + block])
    assert int((f[:, Lp + block:] != 126336).sum()) == 0      # stopped after the first block
    with pytest.raises(NotImplementedError):
        mmada_b200.generate(m, prompt, steps=steps, gen_length=gen, block_length=block, remasking="nope")
    with pytest.raises(AssertionError):
        mmada_b200.generate(m, prompt, steps=steps, gen_length=gen, block_length=7)


def test_generate_random_remasking_matches_oracle_replay(golden):
    """remasking='random' (generate.py:89-90): the confidences are (B, L) fp32 uniforms from the device generator, drawn
    exactly like the reference draws them; decisions replayed on the CPU with the same draws."""
    import mmada_b200
    from oracle import denoise, weights as W
    gd = golden("text_t0")
    B, Lp, gen, block, steps, wseed, seed = (int(v) for v in gd["meta"])
    m = _model(W.TINY, wseed)
    prompt = torch.from_numpy(gd["prompt"])
    L = Lp + gen
    torch.cuda.manual_seed(1234)
    trace = []
    x = mmada_b200.generate(m, prompt.cuda(), steps=steps, gen_length=gen, block_length=block, temperature=0.0,
                            remasking="random", trace=trace)
    # the same stream again: one (B, L) draw per forward, in order
    torch.cuda.manual_seed(1234)
    xo = torch.full((B, L), 126336, dtype=torch.long)
    xo[:, :Lp] = prompt
    spb = steps // (gen // block)
    for t in trace:
        nb, i = t["block"], t["step"]
        lo = Lp + nb * block
        r = torch.rand((B, L), device="cuda").cpu()
        assert torch.equal(t["override"].cpu(), r[:, lo:lo + block].double())
        if i == 0:
            ntt = denoise.get_num_transfer_tokens(xo[:, lo:lo + block] == 126336, spb)
        x0, _ = denoise.text_sample_rows(t["logits"].cpu()[:B * block], 0.0, None)
        assert torch.equal(t["x0"].cpu().reshape(-1), x0)
        masked = xo[:, lo:lo + block] == 126336
        conf = torch.where(masked, r[:, lo:lo + block], torch.tensor(-np.inf))
        for j in range(B):
            _, sel = torch.topk(conf[j], k=int(ntt[j, i]))
            xo[j, lo + sel] = x0.view(B, block)[j, sel]
        assert torch.equal(t["x"].cpu(), xo), (nb, i)
    assert torch.equal(x.cpu(), xo) and int((x == 126336).sum()) == 0


def test_generate_reference_rng_stream_matches_oracle_replay(golden):
    """rng='reference': the Gumbel uniforms are the reference's (B, L, V) fp64 draw from the device generator
    (generate.py:14); with the same seed the same draws replayed on the oracle give the same tokens and states."""
    import mmada_b200
    from oracle import denoise, weights as W
    gd = golden("text_t1")
    B, Lp, gen, block, steps, wseed, seed = (int(v) for v in gd["meta"])
    T = float(gd["temperature"])
    assert T > 0
    m = _model(W.TINY, wseed)
    prompt = torch.from_numpy(gd["prompt"])
    L, V = Lp + gen, W.TINY["vocab_size"]
    torch.cuda.manual_seed(77)
    trace = []
    x = mmada_b200.generate(m, prompt.cuda(), steps=steps, gen_length=gen, block_length=block, temperature=T, rng="reference",
                            trace=trace)
    torch.cuda.manual_seed(77)
    xo = torch.full((B, L), 126336, dtype=torch.long)
    xo[:, :Lp] = prompt
    spb = steps // (gen // block)
    for t in trace:
        nb, i = t["block"], t["step"]
        lo = Lp + nb * block
        u = torch.rand((B, L, V), dtype=torch.float64, device="cuda")[:, lo:lo + block].reshape(B * block, V).cpu()
        assert torch.equal(t["u"].cpu(), u)
        if i == 0:
            ntt = denoise.get_num_transfer_tokens(xo[:, lo:lo + block] == 126336, spb)
        x0, p = denoise.text_sample_rows(t["logits"].cpu()[:B * block], T, u)
        assert torch.equal(t["x0"].cpu().reshape(-1), x0)
        masked = xo[:, lo:lo + block] == 126336
        conf = torch.where(masked, p.view(B, block), torch.tensor(-np.inf, dtype=torch.float64))
        for j in range(B):
            _, sel = torch.topk(conf[j], k=int(ntt[j, i]))
            xo[j, lo + sel] = x0.view(B, block)[j, sel]
        assert torch.equal(t["x"].cpu(), xo), (nb, i)
    assert torch.equal(x.cpu(), xo) and int((x == 126336).sum()) == 0
    with pytest.raises(ValueError):
        mmada_b200.generate(m, prompt.cuda(), steps=steps, gen_length=gen, block_length=block, rng="mt19937")

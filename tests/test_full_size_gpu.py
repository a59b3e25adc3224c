"""Config-2 size (MMaDA-8B architecture, 32 layers, L = 1539, 1024 image tokens, CFG): the CPU oracle cannot run this
in test time, so parity is checked through properties that do not depend on size:
  * determinism: the same prompts and per-prompt noise streams give the same ids twice;
  * the work the kernels skip changes nothing: last block / ln_f / head on the still-masked rows only == on every row;
  * a prompt's result does not depend on what else is in the batch (the property the multi-GPU prompt sharding relies on:
    GEMM rows, attention (batch, head) items and sampling rows are independent, and noise is drawn per prompt);
  * the loop's invariants: ids inside the codebook, input_ids mutated in place to the returned ids + text vocabulary.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

C2 = dict(d_model=4096, n_heads=32, n_layers=32, mlp_hidden_size=12288, vocab_size=134656, rope_theta=500000.0,
          rms_norm_eps=1e-5, max_sequence_length=4096, mask_token_id=126336)
N_IMG, PREFIX, STEPS, GUIDANCE, TEXT_VOCAB = 1024, 513, 4, 3.5, 126349


@pytest.fixture(scope="module")
def model8b():
    from mmada_b200 import MMadaConfig, MMadaModelLM
    if torch.cuda.mem_get_info()[0] < 40 << 30:
        pytest.skip("needs ~20 GB of free device memory")
    m = MMadaModelLM(MMadaConfig.from_dict(C2), device="cuda").init_random(seed=1234)
    yield m
    del m
    torch.cuda.empty_cache()


def _run(model, cond, unc, prompt_ids):
    from mmada_b200.dist import prompt_seed
    from mmada_b200.prompting import UniPromptingLike
    gens = [torch.Generator(device="cuda").manual_seed(prompt_seed(99, i)) for i in prompt_ids]
    ids = cond[prompt_ids].clone().cuda()
    out = model.t2i_generate(input_ids=ids, uncond_input_ids=unc[prompt_ids].cuda(), guidance_scale=GUIDANCE, timesteps=STEPS,
                             seq_len=N_IMG, resolution=PREFIX - 1, generator=gens, uni_prompting=UniPromptingLike(TEXT_VOCAB))
    torch.cuda.synchronize()
    return out.cpu(), ids.cpu()


def test_config2_properties(model8b):
    from mmada_b200.prompting import synthetic_t2i_batch
    cond, unc, _, _ = synthetic_t2i_batch(2, PREFIX, N_IMG, seed=5)
    assert cond.shape[1] == PREFIX + 1 + N_IMG + 1
    m = model8b
    out, ids = _run(m, cond, unc, [0, 1])
    # invariants of the loop (modeling_mmada.py:200-209)
    assert out.shape == (2, N_IMG) and out.dtype == torch.int64
    assert int(out.min()) >= 0 and int(out.max()) < 8192
    img = ids[:, PREFIX + 1:PREFIX + 1 + N_IMG]
    still = img == C2["mask_token_id"]
    assert torch.equal(img[~still], out[~still] + TEXT_VOCAB)
    assert torch.equal(ids[:, :PREFIX + 1], cond[:, :PREFIX + 1])
    # determinism
    out2, ids2 = _run(m, cond, unc, [0, 1])
    assert torch.equal(out, out2) and torch.equal(ids, ids2)
    # the skipped work changes nothing
    m.masked_rows_only, m.restrict_last_block = False, False
    try:
        out3, ids3 = _run(m, cond, unc, [0, 1])
    finally:
        m.masked_rows_only, m.restrict_last_block = True, True
    assert torch.equal(out, out3) and torch.equal(ids, ids3)
    # batch composition / sharding independence
    for i in (0, 1):
        o1, i1 = _run(m, cond, unc, [i])
        assert torch.equal(o1[0], out[i]) and torch.equal(i1[0], ids[i])


def test_config4_text_generation_properties(model8b):
    """generate() (low-confidence remasking, greedy, with and without CFG) at the 8B architecture: deterministic,
    unchanged by the last-block row restriction, and independent of the batch a prompt is in."""
    from mmada_b200 import generate
    g = torch.Generator().manual_seed(7)
    prompt = torch.randint(0, 126000, (2, 96), generator=g)
    kw = dict(steps=8, gen_length=64, block_length=32, temperature=0.0, remasking="low_confidence")
    m = model8b
    for cfg_scale in (0.0, 1.5):
        a = generate(m, prompt.cuda(), cfg_scale=cfg_scale, **kw).cpu()
        assert a.shape == (2, 96 + 64) and torch.equal(a[:, :96], prompt)
        assert not bool((a == C2["mask_token_id"]).any())
        assert torch.equal(a, generate(m, prompt.cuda(), cfg_scale=cfg_scale, **kw).cpu())
        m.restrict_last_block = False
        try:
            b = generate(m, prompt.cuda(), cfg_scale=cfg_scale, **kw).cpu()
        finally:
            m.restrict_last_block = True
        assert torch.equal(a, b)
        for i in (0, 1):
            assert torch.equal(generate(m, prompt[i:i + 1].cuda(), cfg_scale=cfg_scale, **kw).cpu()[0], a[i])

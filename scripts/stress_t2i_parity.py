"""One-off stress run (not part of the suite): whole t2i_generate loops on small head_dim-128 models with random batch sizes,
prompt / image lengths, step counts, guidance scales and noise seeds — every step's sampled ids and masks replayed on the CPU
oracle from the CUDA path's own logits must be bit-identical, the first-step logits within 2e-2 of the oracle's fp32 forward.
usage: python scripts/stress_t2i_parity.py [n_cases]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmada_b200 import MMadaConfig, MMadaModelLM
from mmada_b200.prompting import UniPromptingLike
from oracle import denoise, llada, weights as W

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 12
cfg = W.TINY128
bad = 0
worst = 0.0
for case in range(n_cases):
    g = torch.Generator().manual_seed(500 + case)
    sd = W.make_llada_weights(cfg, case % 3)
    model = MMadaModelLM(MMadaConfig.from_dict(cfg), device="cuda:0").load_state_dict(sd)
    B = int(torch.randint(1, 4, (1,), generator=g))
    P = int(torch.randint(9, 60, (1,), generator=g))
    N = [64, 100, 256, 300, 17][case % 5]
    steps = int(torch.randint(2, 9, (1,), generator=g))
    guidance = [3.5, 0.0, 1.7][case % 3]
    cond, unc, _, _ = W.make_t2i_prompts(B, P, N, case)
    noise = [(torch.empty(B * N, W.CODEBOOK).exponential_(1, generator=g), torch.rand(B, N, generator=g)) for _ in range(steps)]
    ids = cond.clone().cuda()
    trace = []
    out = model.t2i_generate(input_ids=ids, uncond_input_ids=unc.cuda() if guidance > 0 else None, guidance_scale=guidance,
                             timesteps=steps, seq_len=N, resolution=P - 1, uni_prompting=UniPromptingLike(W.TEXT_VOCAB),
                             noise=noise, trace=trace)
    torch.cuda.synchronize()
    known = torch.full((B, N), cfg["mask_token_id"], dtype=torch.int64)
    temperature, ok = 1.0, True
    sched = denoise.t2i_mask_len_schedule(N, steps)
    for s, t in enumerate(trace):
        temperature *= 1.0 - (s + 1) / steps
        r = denoise.t2i_sample_step(t["cond"].cpu(), None if t["uncond"] is None else t["uncond"].cpu(), guidance, known,
                                    cfg["mask_token_id"], sched[s], temperature, *noise[s])
        ok = ok and torch.equal(t["sampled_ids"].cpu(), r["sampled_ids"]) and torch.equal(t["masking"].cpu(), r["masking"])
        known = r["next_known"]
    ok = ok and torch.equal(out.cpu(), r["sampled_ids"])
    rows, cols = slice(-(N + 1), -1), slice(W.TEXT_VOCAB, W.TEXT_VOCAB + W.CODEBOOK)
    ref = llada.forward_logits(cond, sd, cfg, rows, cols)
    err = float((trace[0]["cond"].cpu() - ref).abs().max() / ref.abs().max())
    worst = max(worst, err)
    if not ok or err >= 2e-2:
        bad += 1
        print(f"case {case}: B={B} P={P} N={N} steps={steps} g={guidance}: decisions {'ok' if ok else 'DIFFER'}, logits err {err:.2e}")
print(f"{n_cases} generations, {bad} failures, worst first-step logits error {worst:.2e}")
sys.exit(1 if bad else 0)

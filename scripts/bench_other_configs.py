"""Throughput of the BASELINE.json configurations other than the headline one, through the public API on one GPU
(random-init 8B-architecture weights, synthetic prompts; device-timed with CUDA events after one warm-up run):
  config 3  mmu_generate: 512x512 image (1024 tokens) + question in context, gen 256, block 32, steps 128, B = 1
  config 4  generate(): gen_length 512, block_length 64, steps 256, B = 8 per GPU (64 sharded over 8 GPUs)
  config 5  t2m_generate (B = 1, 16, 64) + motion VQ-VAE decode
  tokenizer side: MAGVITv2.get_code and decode_code at 512x512, B = 8
usage: python scripts/bench_other_configs.py [--what mmu,text,t2m,vq]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def timed(fn, warm=1, reps=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--what", default="mmu,text,t2m,vq")
    args = ap.parse_args()
    what = args.what.split(",")
    from mmada_b200 import MMadaConfig, MMadaModelLM, generate, HumanVQVAE, MAGVITv2
    from mmada_b200.prompting import UniPromptingLike
    import bench
    dev = torch.device("cuda", 0)
    res = {}
    g = torch.Generator().manual_seed(0)
    if any(w in what for w in ("mmu", "text", "t2m")):
        model = MMadaModelLM(MMadaConfig.from_dict(dict(bench.C2)), device=dev).init_random(seed=1234)
    if "mmu" in what:
        # [<|mmu|>, soi, 1024 image tokens, eoi, question (64 tokens)] + 256 masked answer positions
        Lp = 1 + 1 + 1024 + 1 + 64
        idx = torch.cat([torch.tensor([[126089, 126084]]), torch.randint(126349, 126349 + 8192, (1, 1024), generator=g),
                         torch.tensor([[126085]]), torch.randint(0, 126000, (1, 64), generator=g)], 1).to(dev)
        assert idx.shape[1] == Lp
        ms, out = timed(lambda: model.mmu_generate(idx, max_new_tokens=256, steps=128, block_length=32))
        res["config3_mmu_generate"] = dict(ms=ms, ms_per_step=ms / 128, new_tokens_per_s=256 / ms * 1e3, B=1, context=Lp, out_shape=list(out.shape))
    if "text" in what:
        B = 8
        prompt = torch.randint(0, 126000, (B, 64), generator=g).to(dev)
        ms, out = timed(lambda: generate(model, prompt, steps=256, gen_length=512, block_length=64, temperature=1.0, cfg_scale=0.0,
                                         remasking="low_confidence"))
        res["config4_generate"] = dict(ms=ms, ms_per_step=ms / 256, new_tokens_per_s=B * 512 / ms * 1e3, B=B, out_shape=list(out.shape))
    if "t2m" in what:
        up = UniPromptingLike()
        vq = HumanVQVAE(device=dev)
        from oracle import motion                      # synthetic decoder weights only (test infrastructure, not timed)
        vq.load_state_dict(motion.make_motion_decoder_weights(0))
        for B in (1, 16, 64):
            P, N = 257, 256
            ids = torch.cat([torch.randint(0, 126000, (B, P), generator=g), torch.full((B, 1), 126084), torch.full((B, N), 126336),
                             torch.full((B, 1), 126085)], 1).to(dev)
            def run():
                toks = model.t2m_generate(input_ids=ids.clone(), timesteps=15, seq_len=N, uni_prompting=up, temperature=1.0,
                                          image_codebook_size=0)   # the 134656-entry vocabulary has no room behind the image codes
                return vq.forward_decoder_batched(toks.clamp(0, 511))
            ms, out = timed(run)
            res[f"config5_t2m_B{B}"] = dict(ms=ms, motions_per_s=B / ms * 1e3, ms_per_step=ms / 15, out_shape=list(out.shape))
    if "vq" in what:
        vq = MAGVITv2(device=dev).init_random(seed=7)
        sd = {}
        from oracle import weights as W                # synthetic encoder weights only (test infrastructure, not timed)
        vq_e = MAGVITv2(device=dev).load_state_dict({k: v for k, v in W.make_vq_encoder_weights(0).items()})
        B = 8
        px = torch.rand(B, 3, 512, 512, generator=g).mul(2).sub(1).to(dev)
        ms, codes = timed(lambda: vq_e.get_code(px), reps=3)
        res["get_code_512"] = dict(ms=ms, images_per_s=B / ms * 1e3, B=B, out_shape=list(codes.shape))
        ms, pix = timed(lambda: vq.decode_code(codes), reps=3)
        res["decode_code_512"] = dict(ms=ms, images_per_s=B / ms * 1e3, B=B, out_shape=list(pix.shape))
    for k, v in res.items():
        print(k, json.dumps(v))


if __name__ == "__main__":
    main()

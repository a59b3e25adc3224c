"""Where one denoising step's time goes under sustained load: CUDA events around every kernel call of
MMadaModelLM.t2i_generate at BASELINE config 2 (events add ~2 us per call)."""
import collections, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mmada_b200 import MMadaConfig, MMadaModelLM, _lib
from mmada_b200.prompting import UniPromptingLike, synthetic_t2i_batch
import bench

torch.cuda.set_device(0)
model = MMadaModelLM(MMadaConfig.from_dict(bench.C2), device="cuda").init_random(seed=1)
cond, unc, _, _ = synthetic_t2i_batch(8, bench.PREFIX, bench.N_IMG, seed=0)
gen = torch.Generator(device="cuda").manual_seed(1)
def run(stop=None):
    return model.t2i_generate(input_ids=cond.cuda(), uncond_input_ids=unc.cuda(), guidance_scale=3.5, timesteps=15, seq_len=1024,
                              resolution=bench.PREFIX - 1, generator=gen, uni_prompting=UniPromptingLike(), stop_after_steps=stop)
run()                      # warm-up: one full generation
torch.cuda.synchronize()
_lib.EVENTS = []
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run(); e1.record()
torch.cuda.synchronize()
ev, _lib.EVENTS = _lib.EVENTS, None
total = e0.elapsed_time(e1)
agg = collections.OrderedDict()
for name, a, b in ev:
    t, n = agg.get(name, (0.0, 0))
    agg[name] = (t + a.elapsed_time(b), n + 1)
acc = 0.0
print(f"15 steps: {total:.1f} ms  ({total / 15:.2f} ms/step)")
for k, (t, n) in sorted(agg.items(), key=lambda x: -x[1][0]):
    acc += t
    print(f"{t / total * 100:6.2f}%  {t / 15:8.3f} ms/step  {n // 15:4d} calls/step  {t / n * 1e3:9.1f} us/call  {k}")
print(f"{(total - acc) / total * 100:6.2f}%  {(total - acc) / 15:8.3f} ms/step  outside kernel events (torch RNG / copies / gaps)")

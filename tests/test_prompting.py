"""Sequence assembly: the oracle restatement and the device kernel against goldens produced by the
REAL ``UniversalPrompting`` (oracle/make_goldens.py::prompting_case, stub tokenizer)."""
import numpy as np
import pytest
import torch


def _rows(flat, off):
    return [flat[off[i]:off[i + 1]].tolist() for i in range(len(off) - 1)]


def test_oracle_prompt_layouts_match_reference_golden(golden):
    from oracle import prompting as OP
    g = golden("prompting")
    for name in ("t2i_a", "t2i_b"):
        texts = _rows(g[f"{name}_text"], g[f"{name}_off"])
        mtl = int(g[f"{name}_max_text_len"])
        for i, t in enumerate(texts):
            ids, mask = OP.prefix_layout(t, g[f"{name}_image"][i].tolist(), mtl, OP.RESERVED["<|t2i|>"],
                                         OP.RESERVED["<|soi|>"], OP.RESERVED["<|eoi|>"])
            assert ids == g[f"{name}_ids"][i].tolist() and mask == g[f"{name}_mask"][i].tolist()
            ids, mask = OP.prefix_layout(t, g[f"{name}_motion"][i].tolist(), mtl, OP.RESERVED["<|t2m|>"],
                                         OP.RESERVED["<|som|>"], OP.RESERVED["<|eom|>"])
            assert ids == g[f"{name}_t2m_ids"][i].tolist() and mask == g[f"{name}_t2m_mask"][i].tolist()
    texts = _rows(g["mmu_text"], g["mmu_off"])
    for i, t in enumerate(texts):
        seq, plen = OP.mmu_gen_layout(t, g["mmu_image"][i].tolist(), int(g["mmu_max_text_len"]), int(g["end_header"]))
        assert seq == g["mmu_ids"][i].tolist() and plen == int(g["mmu_prompt_length"][i])


def test_host_helpers_match_reference_golden(golden):
    """mmada_b200.prompting's host-side helpers (synthetic batches for the benchmark) produce the reference's layout."""
    from mmada_b200 import prompting as P
    g = golden("prompting")
    for name in ("t2i_a", "t2i_b"):
        texts = _rows(g[f"{name}_text"], g[f"{name}_off"])
        ids, mask = P.t2i_gen_prompt(texts, torch.from_numpy(g[f"{name}_image"]), int(g[f"{name}_max_text_len"]) + 1)
        assert np.array_equal(ids.numpy(), g[f"{name}_ids"]) and np.array_equal(mask.numpy(), g[f"{name}_mask"])
    # the benchmark's synthetic prompts have the length the real class gives for max_text_len = 512
    cond, unc, _, _ = P.synthetic_t2i_batch(2, 513, 1024, seed=0)
    assert cond.shape[1] == g["t2i_b_ids"].shape[1] == 1539


@pytest.mark.gpu
def test_device_prompt_assembly_bit_exact(golden):
    from mmada_b200.prompting import UniversalPrompting
    g = golden("prompting")
    for name in ("t2i_a", "t2i_b"):
        up = UniversalPrompting(None, max_text_len=int(g[f"{name}_max_text_len"]), device="cuda")
        texts = _rows(g[f"{name}_text"], g[f"{name}_off"])
        ids, mask = up.t2i_gen_prompt(texts, torch.from_numpy(g[f"{name}_image"]).cuda())
        assert np.array_equal(ids.cpu().numpy(), g[f"{name}_ids"]) and np.array_equal(mask.cpu().numpy(), g[f"{name}_mask"])
        ids, mask = up.t2m_gen_prompt(texts, torch.from_numpy(g[f"{name}_motion"]).cuda())
        assert np.array_equal(ids.cpu().numpy(), g[f"{name}_t2m_ids"]) and np.array_equal(mask.cpu().numpy(), g[f"{name}_t2m_mask"])
        # a strided view of the image tokens (row pitch != N)
        wide = torch.zeros((len(texts), g[f"{name}_image"].shape[1] + 5), dtype=torch.int64, device="cuda")
        wide[:, 3:-2] = torch.from_numpy(g[f"{name}_image"]).cuda()
        ids2, _ = up.t2i_gen_prompt(texts, wide[:, 3:-2])
        assert np.array_equal(ids2.cpu().numpy(), g[f"{name}_ids"])
    up = UniversalPrompting(None, max_text_len=int(g["mmu_max_text_len"]), device="cuda", end_header_id=int(g["end_header"]))
    texts = _rows(g["mmu_text"], g["mmu_off"])
    ids, plen = up.mmu_gen_prompt(torch.from_numpy(g["mmu_image"]).cuda(), texts)
    assert np.array_equal(ids.cpu().numpy(), g["mmu_ids"]) and np.array_equal(plen.cpu().numpy(), g["mmu_prompt_length"])
    # inference_mmu.py:93-100
    from oracle import prompting as OP
    img = torch.from_numpy(g["mmu_image"]).cuda()
    q = [5, 6, 7]
    got = up.mmu_input_ids(img, q).cpu()
    for i in range(img.shape[0]):
        assert got[i].tolist() == OP.mmu_inference_layout(q, g["mmu_image"][i].tolist())


@pytest.mark.gpu
def test_uint8_postprocess_vs_reference_decode_golden(golden):
    """inference_t2i.py:123-126 (clamp((x+1)/2, 0, 1) * 255 -> uint8, NHWC) on the CUDA decoder's output against the same
    post-process of the reference's fp32 ``decode_code`` golden: the difference a user sees in the saved image."""
    from mmada_b200.modeling_magvitv2 import MAGVITv2
    from oracle import weights as W
    g = golden("magvit")
    vq = MAGVITv2(device="cuda").load_state_dict(W.make_vq_decoder_weights(0))
    d_all = []
    for tag in ("16x16",):                        # (the 8x8 golden is below the conv kernel's smallest tile)
        codes = torch.from_numpy(g[f"idx_{tag}"]).cuda()
        mine = vq.decode_code_uint8(codes).cpu().numpy().astype(np.int32)[:, ::2, ::2, :]    # the golden keeps every 2nd pixel
        ref = torch.from_numpy(g[f"pix_{tag}"]).float()
        ref8 = (torch.clamp((ref + 1.0) / 2.0, 0.0, 1.0) * 255.0).permute(0, 2, 3, 1).numpy().astype(np.uint8).astype(np.int32)
        d_all.append(np.abs(mine - ref8).reshape(-1))
    d = np.concatenate(d_all)
    mine = ref8 = None
    print(f"uint8 image vs the reference's fp32 decode: max |d| = {int(d.max())}, mean |d| = {d.mean():.3f} (of 255)")
    assert d.max() <= 12 and d.mean() <= 1.5

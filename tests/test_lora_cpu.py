"""CPU: the LoRA-merge oracle against the fixture produced by the reference's own lora.py (oracle/make_golden_lora.py),
and the product's key handling against both.  No compute through the product here (it has no CPU path)."""
import pytest
import torch

import ltx_oracle as O
import mlx_video_b200 as M
from mlx_video_b200 import lora


def lora_fixture(golden):
    g = golden("lora")
    base = {k[5:]: (torch.from_numpy(g[k]).to(torch.bfloat16) if g[k].ndim == 2 else torch.from_numpy(g[k]))
            for k in g if k.startswith("base/")}
    loras = []
    for n, s in enumerate(g["strengths"]):
        sd = {k.split("/", 1)[1]: torch.from_numpy(g[k]) for k in g if k.startswith(f"lora{n}/")}
        loras.append((sd, float(s)))
    merged = {k[7:]: torch.from_numpy(g[k]) for k in g if k.startswith("merged/")}
    return base, loras, merged


def test_oracle_merge_is_bit_identical_to_the_reference(golden):
    base, loras, merged = lora_fixture(golden)
    got = O.apply_lora_to_weights(base, loras)
    assert set(got) == set(merged)
    changed = 0
    for k, want in merged.items():
        assert got[k].dtype == base[k].dtype
        assert torch.equal(got[k].float(), want), k
        changed += int(not torch.equal(got[k], base[k]))
    assert changed == 5  # the pair whose base weight does not exist is skipped, the bias is untouched


def test_product_key_handling_matches_reference_names(golden):
    base, loras, _ = lora_fixture(golden)
    for sd, _ in loras:
        ours = [(r, s, tuple(a.shape), tuple(b.shape)) for r, s, a, b in lora._iter_lora_pairs(sd)]
        want = [(r, s, tuple(a.shape), tuple(b.shape)) for r, s, a, b in O.lora_pairs(sd)]
        assert ours == want and len(ours) >= 4
        for raw, san, _, _ in ours:
            assert lora._candidate_weight_keys(raw, san) == tuple(O.lora_candidate_keys(raw, san))
            hit = next((k for k in lora._candidate_weight_keys(raw, san) if k in base), None)
            assert (hit is None) == ("not_in_model" in raw)
    # the renames of LTXModel.sanitize (lora.py:18-33)
    f = lora._sanitize_lora_prefix
    assert f("model.diffusion_model.transformer_blocks.3.attn1.to_out.0.weight") == "transformer_blocks.3.attn1.to_out.weight"
    assert f("diffusion_model.transformer_blocks.3.ff.net.0.proj.weight") == "transformer_blocks.3.ff.proj_in.weight"
    assert f("transformer_blocks.3.audio_ff.net.2.weight") == "transformer_blocks.3.audio_ff.proj_out.weight"
    assert f("adaln_single.emb.timestep_embedder.linear_1.weight") == "adaln_single.emb.timestep_embedder.linear1.weight"
    assert lora.has_quantized_weights({"a.scales": 0}) and not lora.has_quantized_weights({"a.weight": 0})


def test_merge_has_no_cpu_path():
    w = torch.zeros(64, 64, dtype=torch.bfloat16)
    with pytest.raises(M.LtxbError):
        lora.merge_lora_pair(w, torch.zeros(8, 64), torch.zeros(64, 8), 1.0)

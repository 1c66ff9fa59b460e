"""GPU: LoRA merge through the C ABI (ltxb_gemm_bf16 + ltxb_lora_merge_bf16) against the reference's own result
(tests/golden/lora.npz) and, end to end, a model with a merged LoRA against the oracle on oracle-merged weights."""
import numpy as np
import pytest
import torch
from safetensors.torch import save_file

import ltx_oracle as O
import mlx_video_b200 as M
from conftest import cosine, rel_l2
from mlx_video_b200 import lora, sampler
from test_lora_cpu import lora_fixture

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def write_loras(tmp_path, loras):
    specs = []
    for n, (sd, s) in enumerate(loras):
        p = tmp_path / f"lora{n}.safetensors"
        save_file({k: v.contiguous() for k, v in sd.items()}, str(p))
        specs.append(lora.LoraSpec(p, s))
    return specs


def test_merge_matches_reference_golden(golden, tmp_path):
    """bf16 weights: identical to the reference's merge except where the fp32 delta sits on a bf16 rounding boundary
    (the tensor-core sum runs in another order): at most 1 bf16 ulp (of the largest of w, delta, w + delta), on fewer
    than 0.2 % of the elements."""
    base, loras, merged = lora_fixture(golden)
    out = lora.apply_lora_to_weights(base, write_loras(tmp_path, loras), device=DEV)
    assert set(out) == set(base)
    for k, want in merged.items():
        got = out[k]
        if torch.equal(want, base[k].float()):
            assert got is base[k], f"{k}: untouched entries must be passed through"
            continue
        assert got.is_cuda and got.dtype == torch.bfloat16 and not torch.equal(base[k], got.cpu())
        g32, w32 = got.float().cpu(), want
        # a flip happens in bf16(delta_n) or in the rounding of a running sum: 1 ulp of the largest value on the way
        run = base[k].float()
        mag = run.abs()
        for sd, strength in loras:
            for raw, san, A, B in O.lora_pairs(sd):
                if san == k:
                    delta = (B.float() @ A.float()) * strength
                    run = (run.to(torch.bfloat16) + delta.to(torch.bfloat16)).float()
                    mag = torch.maximum(mag, torch.maximum(delta.abs(), run.abs()))
        assert torch.equal(run, w32), "the test's own restatement of the merge drifted from the golden"
        ulp = torch.maximum(mag, torch.tensor(1e-30)).log2().floor().exp2() * 2.0 ** -7
        bad = (g32 - w32).abs() > 0
        assert float(((g32 - w32).abs() / ulp).max()) <= 1.0 + 1e-6, k
        assert float(bad.float().mean()) < 2e-3, f"{k}: {float(bad.float().mean()):.2e} of the elements differ"
    for k in base:  # the caller's tensors are never written
        assert torch.equal(base[k], lora_fixture(golden)[0][k])


def test_model_with_merged_lora_vs_oracle(tmp_path):
    """apply_lora_to_model on a 2-block model — pairs on fused q|k|v rows, on the stacked text K/V and on the FFN —
    against the fp32 oracle running on oracle-merged weights: rel-L2 <= 1e-2, cosine >= 0.999; and the LoRA must
    actually move the output."""
    cfg = O.small_config(O.LTXModelType.VideoOnly, num_layers=2)
    tensors = {k: (v.to(torch.bfloat16) if k.endswith(".weight") and v.dim() == 2 else v) for k, v in O.init_params(cfg, seed=3).items()}
    g = torch.Generator().manual_seed(11)
    D = cfg.num_attention_heads * cfg.attention_head_dim
    sd = {}
    for prefix, o, i in [("diffusion_model.transformer_blocks.0.attn1.to_k", D, D), ("diffusion_model.transformer_blocks.1.attn1.to_out.0", D, D),
                         ("diffusion_model.transformer_blocks.1.attn2.to_v", D, D), ("diffusion_model.transformer_blocks.0.ff.net.0.proj", 4 * D, D),
                         ("diffusion_model.transformer_blocks.1.ff.net.2", D, 4 * D)]:
        sd[f"{prefix}.lora_A.weight"] = (0.2 * torch.randn(24, i, generator=g)).to(torch.bfloat16)
        sd[f"{prefix}.lora_B.weight"] = (0.2 * torch.randn(o, 24, generator=g)).to(torch.bfloat16)
    specs = write_loras(tmp_path, [(sd, 0.7)])
    merged = O.apply_lora_to_weights(tensors, [(sd, 0.7)])
    d = {k: getattr(cfg, k) for k in cfg.__dataclass_fields__}
    d["model_type"], d["rope_type"] = d["model_type"].value, d["rope_type"].value
    model = M.LTXModel(M.LTXModelConfig.from_dict(d), device=DEV)
    model.load_weights({k: v.float() for k, v in tensors.items()})
    F_, H_, W_, Tc = 3, 8, 8, 32
    T = F_ * H_ * W_
    lat, ctx = torch.randn(1, T, 128, generator=g), torch.randn(1, Tc, 256, generator=g)
    pos = torch.from_numpy(sampler.create_position_grid(1, F_, H_, W_))
    ts = torch.full((1, T), 0.421875)
    mod = M.Modality(lat.to(DEV), ts.to(DEV), pos.to(DEV), ctx.to(DEV))
    before, _ = model(video=mod, audio=None)
    before = before.float().cpu()
    lora.apply_lora_to_model(model, specs)
    got, _ = model(video=mod, audio=None)
    got = got.float().cpu()
    want, _ = O.OracleLTXModel(cfg, {k: v.float() for k, v in merged.items()})(O.Modality(lat, ts, pos, ctx), None)
    r, c = rel_l2(got, want), cosine(got, want)
    assert r <= 1e-2 and c >= 0.999, f"merged-LoRA forward: rel_l2={r:.3e} cos={c:.6f}"
    assert rel_l2(before, want) > 5 * r, "the LoRA did not change the output enough to test anything"

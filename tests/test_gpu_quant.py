"""GPU: MLX affine-quantised checkpoints through the C ABI (ltxb_dequant_affine_bf16 + the bf16 GEMMs) — the expanded
weights bit-exact against the oracle's dequantisation, and ``LTXModel.from_pretrained`` + forward against the
velocities the reference's own ``from_pretrained`` + forward produced (tests/golden/quant.npz)."""
import numpy as np
import pytest
import torch

import ltx_oracle as O
import mlx_video_b200 as M
import quant_fixture as QF
from conftest import rel_l2
from mlx_video_b200 import lora, ops
from test_gpu_parity import assert_close, product_config, to_dev

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("aux", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("bits", [2, 4, 8])
@pytest.mark.parametrize("group_size", [32, 64, 128])
def test_dequant_kernel_is_bit_exact(bits, group_size, aux):
    """out = bf16(scales * q + biases); also into a row-strided view (fused q|k|v storage) and with f32 scales."""
    g = torch.Generator().manual_seed(7 * bits + group_size)
    R, C = 200, 384
    w = (torch.randn(R, C, generator=g) / 8).to(aux)
    packed, s, b = O.affine_quantize(w, group_size, bits)
    want = O.affine_dequantize(packed, s, b, group_size, bits).to(torch.bfloat16)
    store = torch.full((R, 3 * C), 7.0, dtype=torch.bfloat16, device=DEV)
    out = store[:, C:2 * C]
    ops.dequant_affine(torch.from_numpy(packed.view(np.int32)).to(DEV), s.to(DEV), b.to(DEV), out, group_size, bits)
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), want), f"max |diff| {float((out.cpu().float() - want.float()).abs().max())}"
    assert bool((store[:, :C] == 7).all()) and bool((store[:, 2 * C:] == 7).all()), "wrote outside its column range"


def test_dequant_argument_errors():
    out = torch.empty(8, 256, dtype=torch.bfloat16, device=DEV)
    p = torch.zeros(8, 32, dtype=torch.int32, device=DEV)
    s = torch.zeros(8, 4, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(ValueError):
        ops.dequant_affine(p, s, s, out, 64, 8)  # 32 words per row are 4-bit levels of a 256-wide row, not 8-bit
    with pytest.raises(M.LtxbError):
        ops.dequant_affine(p.cpu(), s, s, out, 64, 4)  # no CPU path
    with pytest.raises(M.LtxbError, match="bits"):
        ops.dequant_affine(torch.zeros(8, 24, dtype=torch.int32, device=DEV), s, s, out, 64, 3)


@pytest.mark.parametrize("variant", sorted(QF.VARIANTS))
def test_from_pretrained_quantised_checkpoint(golden, tmp_path, variant):
    g = golden("quant")
    cfg = QF.config()
    path = QF.write_checkpoint(variant, tmp_path)
    state, dense = QF.build(variant)
    assert QF.packed_checksum(state) == int(g[f"{variant}/checksum"][0])
    model = M.LTXModel.from_pretrained(path, product_config(cfg), strict=True, device=DEV)
    params = model.parameters()
    n_q = 0
    for k, want in dense.items():  # the device weights are exactly the bf16 expansion of the packed tensors
        if isinstance(state[k], np.ndarray):
            assert torch.equal(params[k].cpu(), want.to(torch.bfloat16)), k
            n_q += 1
    assert n_q == int(g[f"{variant}/n_quantized"][0])
    got, none = model(video=to_dev(QF.inputs(variant)), audio=None)
    torch.cuda.synchronize()
    assert none is None
    assert_close(got.cpu(), torch.from_numpy(g[f"{variant}/velocity"]), f"quantised checkpoint ({variant})")
    # same result from tensors already in memory (weights_override, ltx.py:617-623)
    in_memory = {k: (torch.from_numpy(v.view(np.int32)) if isinstance(v, np.ndarray) else v) for k, v in state.items()}
    model2 = M.LTXModel.from_pretrained(None, product_config(cfg), strict=True, weights_override=in_memory, device=DEV)
    got2, _ = model2(video=to_dev(QF.inputs(variant)), audio=None)
    assert torch.equal(got2, got)
    # LoRA on top: the reference attaches runtime adapters to a quantised model (lora.py:219-275); here the same call
    # merges into the expanded weights — compared with the velocity the reference's adapters produced
    spec = lora.LoraSpec(QF.write_lora(variant, tmp_path), QF.LORA_STRENGTH)
    lora.apply_lora_to_model(model, [spec])
    got_l, _ = model(video=to_dev(QF.inputs(variant)), audio=None)
    torch.cuda.synchronize()
    assert_close(got_l.cpu(), torch.from_numpy(g[f"{variant}/velocity_lora"]), f"quantised checkpoint + LoRA ({variant})")
    assert rel_l2(got_l.float(), got.float()) > 2e-2, "the LoRA did not change the output"
    with pytest.raises(ValueError, match="quantised"):
        lora.apply_lora_to_weights(in_memory, [spec], device=DEV)


def test_from_pretrained_errors(tmp_path):
    cfg = QF.config()
    state, _ = QF.build("mlx")
    entries = {k: QF._entry(v) for k, v in state.items() if k != "transformer_blocks.1.ff.proj_out.bias"}
    QF.write_safetensors(tmp_path / "broken.safetensors", entries)
    with pytest.raises(ValueError, match="Missing 1 parameters after load"):  # ltx.py:874-881
        M.LTXModel.from_pretrained(tmp_path / "broken.safetensors", product_config(cfg), strict=True, device=DEV)
    M.LTXModel.from_pretrained(tmp_path / "broken.safetensors", product_config(cfg), strict=False, device=DEV)
    d = tmp_path / "meta"
    path = QF.write_checkpoint("upstream", d)
    (d / "quantization.json").write_text('{"group_size": 64, "bits": 4}')
    with pytest.raises(ValueError, match="quantization.json"):
        M.LTXModel.from_pretrained(path, product_config(cfg), strict=True, device=DEV)
    entries = {k: QF._entry(v) for k, v in state.items() if not k.endswith("attn1.to_q.biases")}
    QF.write_safetensors(tmp_path / "nobias.safetensors", entries)
    with pytest.raises(ValueError, match="biases"):
        M.LTXModel.from_pretrained(tmp_path / "nobias.safetensors", product_config(cfg), strict=True, device=DEV)

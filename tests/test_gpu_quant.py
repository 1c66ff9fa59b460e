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


def _quantised(N, K, group_size, bits, aux, seed):
    g = torch.Generator().manual_seed(seed)
    w = (torch.randn(N, K, generator=g) / (K ** 0.5)).to(aux)
    packed, s, b = O.affine_quantize(w, group_size, bits)
    return torch.from_numpy(packed.view(np.int32)).to(DEV), s.to(DEV), b.to(DEV)


# (M, N, K, group, bits): sequence-parallel shards, audio tokens, AdaLN rows, ragged N, the largest M, one k-block
QW_SHAPES = [(160, 4096, 4096, 64, 4), (160, 4096, 4096, 64, 8), (160, 12288, 4096, 128, 4), (160, 4096, 16384, 64, 4),
             (68, 2048, 2048, 32, 4), (68, 2048, 2048, 32, 8), (1, 1024, 512, 64, 8), (200, 272, 1024, 128, 8),
             (256, 4096, 4096, 64, 4), (250, 784, 256, 32, 4), (16, 256, 64, 64, 4)]


@pytest.mark.parametrize("aux", [torch.bfloat16, torch.float32])
@pytest.mark.parametrize("M_,N,K,group_size,bits", QW_SHAPES)
def test_packed_weight_gemm_is_bit_identical_to_dequant_then_gemm(M_, N, K, group_size, bits, aux):
    """ltxb_gemm_qw_bf16 (packed tiles from HBM, expanded in shared memory) == ltxb_dequant_affine_bf16 + the few-row bf16 GEMM
    with the same k-range splits, bit for bit; and both agree with an fp32 matmul over the dequantised weight."""
    packed, s, b = _quantised(N, K, group_size, bits, aux, 11 * bits + group_size + M_)
    g = torch.Generator(device=DEV).manual_seed(M_ + N)
    a = torch.randn(M_, K, device=DEV, generator=g).bfloat16()
    bias = torch.randn(N, device=DEV, generator=g)
    w = torch.empty(N, K, dtype=torch.bfloat16, device=DEV)
    ops.dequant_affine(packed, s, b, w, group_size, bits)
    ref = a.float() @ w.float().T + bias
    tiles = (N + 255) // 256
    for splits in (0, 1, 2, 3):
        want = torch.empty(M_, N, device=DEV, dtype=torch.float32)
        ops.gemm(a, w, bias, want, mode=M._lib.EPI_BIAS_F32, cta_pair=4, block_n=splits)
        got = torch.empty(M_, N, device=DEV, dtype=torch.float32)
        ops.gemm_qw(a, packed, s, b, group_size, bits, bias, got, mode=M._lib.EPI_BIAS_F32, splits=splits)
        again = torch.empty(M_, N, device=DEV, dtype=torch.float32)
        ops.gemm_qw(a, packed, s, b, group_size, bits, bias, again, mode=M._lib.EPI_BIAS_F32, splits=splits, const_w=True)
        torch.cuda.synchronize()
        assert torch.equal(got, again)
        assert rel_l2(got, ref) < 2e-5
        # same k-range splits => same summation order => same bits.  The packed kernel runs one CTA pair per SM pair (74
        # slots), the bf16 one up to two: an explicit split count both can place is honoured by both
        if splits >= 1 and tiles * splits <= 74 and K // 64 >= splits:
            assert torch.equal(got, want), f"splits={splits}: max |diff| {float((got - want).abs().max())}"
        else:
            assert rel_l2(got, want) < 2e-5


@pytest.mark.parametrize("mode", ["bias_bf16", "gelu", "silu", "resid_gate"])
def test_packed_weight_gemm_epilogues(mode):
    E = M._lib
    M_, N, K, group_size, bits = 160, 4096, 4096, 64, 4
    packed, s, b = _quantised(N, K, group_size, bits, torch.bfloat16, 5)
    g = torch.Generator(device=DEV).manual_seed(9)
    a = torch.randn(M_, K, device=DEV, generator=g).bfloat16()
    bias = torch.randn(N, device=DEV, generator=g)
    w = torch.empty(N, K, dtype=torch.bfloat16, device=DEV)
    ops.dequant_affine(packed, s, b, w, group_size, bits)
    if mode == "resid_gate":
        resid = torch.randn(M_, N, device=DEV, generator=g)
        gate = torch.randn(2, N, device=DEV, generator=g)
        table = torch.randn(N, device=DEV, generator=g)
        want, got = resid.clone(), resid.clone()
        ops.gemm(a, w, bias, want, mode=E.EPI_RESID_GATE_F32, resid=want, gate=gate, gate_table=table, gate_row_div=80, cta_pair=4)
        ops.gemm_qw(a, packed, s, b, group_size, bits, bias, got, mode=E.EPI_RESID_GATE_F32, resid=got, gate=gate, gate_table=table, gate_row_div=80)
    else:
        m = {"bias_bf16": E.EPI_BIAS_BF16, "gelu": E.EPI_GELU_BF16, "silu": E.EPI_SILU_BF16}[mode]
        want = torch.empty(M_, N, device=DEV, dtype=torch.bfloat16)
        got = torch.empty_like(want)
        ops.gemm(a, w, bias, want, mode=m, cta_pair=4)
        ops.gemm_qw(a, packed, s, b, group_size, bits, bias, got, mode=m)
    torch.cuda.synchronize()
    assert torch.equal(got, want)


def test_packed_weight_gemm_argument_errors():
    packed, s, b = _quantised(256, 512, 64, 4, torch.bfloat16, 3)
    a = torch.zeros(600, 512, dtype=torch.bfloat16, device=DEV)
    with pytest.raises(M.LtxbError, match="few-row"):
        ops.gemm_qw(a, packed, s, b, 64, 4, None, torch.empty(600, 256, dtype=torch.bfloat16, device=DEV))
    with pytest.raises(M.LtxbError, match="tensor memory"):
        ops.gemm_qw(a[:320], packed, s, b, 64, 4, None, torch.empty(320, 256, dtype=torch.bfloat16, device=DEV))
    p2, s2, b2 = _quantised(256, 512, 64, 2, torch.bfloat16, 3)
    with pytest.raises(M.LtxbError, match="bits"):
        ops.gemm_qw(a[:8], p2, s2, b2, 64, 2, None, torch.empty(8, 256, dtype=torch.bfloat16, device=DEV))
    with pytest.raises(ValueError):
        ops.gemm_qw(a[:8], packed, s, b, 64, 8, None, torch.empty(8, 256, dtype=torch.bfloat16, device=DEV))


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
    # keep_packed: few-row products stream the packed words (ltxb_gemm_qw_bf16) — same bits as the expanded path
    from mlx_video_b200.packed import REGISTRY

    model3 = M.LTXModel.from_pretrained(path, product_config(cfg), strict=True, device=DEV, keep_packed=True)
    assert 0 < len(REGISTRY) <= n_q  # q | k | v of one fused matrix merge into one entry
    launches = M._lib.lib.ltxb_kernel_launches()
    got3, _ = model3(video=to_dev(QF.inputs(variant)), audio=None)
    torch.cuda.synchronize()
    assert torch.equal(got3, got), f"packed-weight path differs: {rel_l2(got3.float(), got.float()):.3e}"
    assert M._lib.lib.ltxb_kernel_launches() > launches
    REGISTRY.clear()
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

"""GPU: the latent upsampler through the C ABI (ltxb_im2col_cl + ltxb_gemm_bf16, ltxb_groupnorm_silu,
ltxb_pixel_shuffle2, ltxb_latent_layout) — every kernel against the oracle's piece, and ``upsample_latents`` /
``load_upsampler`` end to end against the outputs of the reference's own upsampler.py (tests/golden/upsampler.npz).
Tolerance end to end as for the DiT: rel-L2 <= 1e-2, cosine >= 0.999 (bf16 GEMM operands, fp32 everywhere else)."""
import pytest
import torch
import torch.nn.functional as F
from safetensors.torch import save_file

import mlx_video_b200 as M
import upsampler_fixture as UF
import upsampler_oracle as U
from conftest import rel_l2
from mlx_video_b200 import _lib, ops
from test_gpu_parity import assert_close

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("shape,kernel", [((1, 2, 4, 6, 128), (3, 3, 3)), ((2, 1, 3, 5, 64), (3, 3, 3)), ((3, 1, 4, 4, 128), (1, 3, 3))])
def test_im2col_rows_are_the_convolution_operand(shape, kernel):
    """Bit-exact gather: the rows times the flattened (C_out, kd, kh, kw, C_in) weight ARE the convolution."""
    g = torch.Generator().manual_seed(5)
    x = torch.randn(*shape, generator=g)
    N, D, H, W, C = shape
    taps = kernel[0] * kernel[1] * kernel[2]
    cols = torch.empty(N * D * H * W, taps * C, dtype=torch.bfloat16, device=DEV)
    ops.im2col_cl(x.to(DEV), cols, *kernel)
    torch.cuda.synchronize()
    xb = x.to(torch.bfloat16).float()  # what the kernel rounds to
    pad = [kernel[0] // 2, kernel[1] // 2, kernel[2] // 2]
    xp = F.pad(xb, (0, 0, pad[2], pad[2], pad[1], pad[1], pad[0], pad[0]))
    want = torch.stack([xp[:, kz:kz + D, ky:ky + H, kx:kx + W, :] for kz in range(kernel[0]) for ky in range(kernel[1])
                        for kx in range(kernel[2])], dim=4).reshape(N * D * H * W, taps * C)
    assert torch.equal(cols.float().cpu(), want)
    # and through the GEMM: one 3-D convolution against the oracle's
    O = 128
    w = (torch.randn(O, *kernel, C, generator=g) / (taps * C) ** 0.5).to(torch.bfloat16)
    b = torch.randn(O, generator=g)
    out = torch.empty(N * D * H * W, O, dtype=torch.float32, device=DEV)
    ops.gemm(cols, w.reshape(O, -1).to(DEV), b.to(DEV), out, mode=_lib.EPI_BIAS_F32)
    ref = (U.conv3d_cl(xb, w.float(), b) if kernel[0] == 3 else
           U.conv2d_cl(xb.reshape(N * D, H, W, C), w.float()[:, 0], b).reshape(N, D, H, W, O))
    assert rel_l2(out.cpu().reshape(ref.shape), ref) <= 1e-5


@pytest.mark.parametrize("resid,silu", [(False, True), (True, True), (False, False)])
def test_groupnorm_silu_vs_oracle(resid, silu):
    g = torch.Generator().manual_seed(6)
    x = torch.randn(2, 3, 5, 7, 128, generator=g) * 2 + 0.5  # S = 105 rows: a ragged last chunk
    w, b = 1 + 0.1 * torch.randn(128, generator=g), 0.1 * torch.randn(128, generator=g)
    r = torch.randn(x.shape, generator=g) if resid else None
    want = U.group_norm_cl(x, w, b)
    if resid:
        want = want + r
    if silu:
        want = F.silu(want)
    out = torch.empty_like(x, device=DEV)
    ops.groupnorm_silu(x.to(DEV), out, 32, 1e-5, w.to(DEV), b.to(DEV), resid=None if r is None else r.to(DEV), silu=silu)
    torch.cuda.synchronize()
    assert float((out.cpu() - want).abs().max()) <= 2e-5 * float(want.abs().max())


def test_pixel_shuffle_and_layout_kernels():
    g = torch.Generator().manual_seed(7)
    x = torch.randn(3, 4, 5, 4 * 64, generator=g)
    out = torch.empty(3, 8, 10, 64, device=DEV)
    ops.pixel_shuffle2(x.to(DEV), out)
    assert torch.equal(out.cpu(), U.pixel_shuffle_cl(x))
    lat = torch.randn(2, 128, 30, generator=g)
    mean, std = torch.randn(128, generator=g), 0.5 + torch.rand(128, generator=g)
    cl = torch.empty(2, 30, 128, device=DEV)
    ops.latent_layout(lat.to(DEV), cl, std.to(DEV), mean.to(DEV), True)
    assert torch.allclose(cl.cpu(), (lat * std[None, :, None] + mean[None, :, None]).transpose(1, 2), rtol=1e-6, atol=1e-6)
    back = torch.empty(2, 128, 30, device=DEV)
    ops.latent_layout(cl, back, std.to(DEV), mean.to(DEV), False)
    assert torch.allclose(back.cpu(), lat, rtol=1e-5, atol=1e-5)
    plain = torch.empty(2, 30, 128, device=DEV)
    ops.latent_layout(lat.to(DEV), plain, None, None, True)
    assert torch.equal(plain.cpu(), lat.transpose(1, 2))


@pytest.mark.parametrize("case", sorted(UF.CASES))
def test_upsample_latents_against_reference_golden(golden, tmp_path, case):
    g = golden("upsampler")
    c = UF.CASES[case]
    latent, mean, std = UF.inputs(case)
    if case == "loaded":  # upstream conv layouts on disk -> load_upsampler (upsampler.py:319-373)
        path = tmp_path / "upsampler.safetensors"
        state = UF.upstream_state(case)
        state["upsampler.blur_down.kernel"] = torch.ones(1, 1, 5, 5) / 25.0  # carried by upstream files, unused in the forward
        save_file(state, str(path))
        model = M.load_upsampler(path, device=DEV)
        assert model.mid_channels == c["mid"] and len(model.res_blocks) == 4
    else:
        model = M.LatentUpsampler(128, c["mid"], c["blocks"], device=DEV)
        model.load_weights(UF.params(case), strict=True)
    got = M.upsample_latents(latent.to(DEV), model, mean.to(DEV), std.to(DEV))
    torch.cuda.synchronize()
    want = torch.from_numpy(g[f"{case}/output"])
    assert got.shape == want.shape and got.dtype == torch.float32
    assert_close(got.cpu(), want, f"upsample_latents ({case})")
    got16 = M.upsample_latents(latent.to(DEV, torch.bfloat16), model, mean, std)  # bf16 latents in -> bf16 out; CPU statistics are moved
    assert got16.dtype == torch.bfloat16
    assert_close(got16.float().cpu(), want, f"upsample_latents bf16 io ({case})")


def test_upsampler_errors():
    with pytest.raises(ValueError):
        M.LatentUpsampler(128, 96, 1, device=DEV)  # widths the GEMM tiling does not take
    model = M.LatentUpsampler(128, 128, 1, device=DEV)
    with pytest.raises(M.LtxbError):
        model(torch.zeros(1, 128, 1, 2, 2))  # CPU tensor: no fallback
    with pytest.raises(ValueError):
        model(torch.zeros(1, 64, 1, 2, 2, device=DEV))
    with pytest.raises(ValueError, match="missing"):
        model.load_weights({"initial_conv.weight": torch.zeros(128, 3, 3, 3, 128)}, strict=True)

"""CPU: the oracle (oracle/ltx_oracle.py) against the committed golden vectors, which were produced by
running the REFERENCE's own sources over oracle/mlx_shim (oracle/make_golden.py)."""
import numpy as np
import pytest
import torch

import ltx_oracle as O
from conftest import rel_l2

CASES = {
    "video_L2": (O.LTXModelType.VideoOnly, 2),
    "video_L2_b2_i2v": (O.LTXModelType.VideoOnly, 2),
    "video_L1_mask": (O.LTXModelType.VideoOnly, 1),
    "av_L2": (O.LTXModelType.AudioVideo, 2),
    "audio_L1": (O.LTXModelType.AudioOnly, 1),
}


def modalities(g):
    def mod(p):
        if p + "latent" not in g:
            return None
        cm = torch.from_numpy(g[p + "context_mask"]) if p + "context_mask" in g else None
        return O.Modality(torch.from_numpy(g[p + "latent"]), torch.from_numpy(g[p + "timesteps"]),
                          torch.from_numpy(g[p + "positions"]), torch.from_numpy(g[p + "context"]), True, cm)
    return mod("v_"), mod("a_")


def test_grids_bit_exact(golden):
    g = golden("grids")
    assert np.array_equal(O.create_position_grid(1, 5, 16, 16), g["g_1_5_16_16"])
    assert np.array_equal(O.create_position_grid(2, 3, 4, 6), g["g_2_3_4_6"])
    assert np.array_equal(O.create_position_grid(1, 9, 24, 24)[:, :, ::7, :], g["g_1_9_24_24"])
    assert np.array_equal(O.create_audio_position_grid(1, 68), g["a_1_68"])
    assert np.array_equal(O.create_audio_position_grid(2, 21), g["a_2_21"])
    for steps, ntok in [(40, 5184), (8, 1280), (30, None), (4, 320)]:
        assert np.array_equal(O.ltx2_scheduler(steps, ntok), g[f"sched_{steps}_{ntok}"])
    assert list(g["STAGE_1_SIGMAS"]) == O.STAGE_1_SIGMAS and list(g["STAGE_2_SIGMAS"]) == O.STAGE_2_SIGMAS


def test_scheduler_properties():
    # the reference's own property tests (tests/test_generate_dev.py:21-70): endpoints, monotone, length
    s = O.ltx2_scheduler(40, 5184)
    assert s.shape == (41,) and s[0] == 1.0 and s[-1] == 0.0 and np.all(np.diff(s) < 0)
    assert abs(float(s[-2]) - 0.1) < 1e-6  # stretched to terminal=0.1


def test_position_grid_properties():
    # tests/test_generate_dev.py:76-142: shape, dtype, causal fix moves the temporal bounds
    g = O.create_position_grid(1, 5, 16, 16)
    assert g.shape == (1, 3, 1280, 2) and g.dtype == np.float32
    assert not np.array_equal(g, O.create_position_grid(1, 5, 16, 16, causal_fix=False))
    assert g[0, 0, 0, 0] == 0.0 and g[0, 0, 0, 1] == np.float32(1 / 24.0)
    assert g[0, 1, 17, 0] == 32.0 and g[0, 2, 17, 0] == 32.0  # token 17 = (f0, h1, w1)


def test_rope_tables(golden):
    g = golden("rope")
    pos = torch.from_numpy(O.create_position_grid(1, 5, 16, 16))
    c, s = O.precompute_freqs_cis(pos, 4096, 10000.0, [20, 2048, 2048], True, 32, O.LTXRopeType.SPLIT, True)
    assert c.shape == (1, 32, 1280, 64) and c.dtype == torch.float32
    assert np.array_equal(c[:, [0, 1, 17, 31], ::5].numpy(), g["video_prod_split_dbl_cos"])
    assert np.array_equal(s[:, [0, 1, 17, 31], ::5].numpy(), g["video_prod_split_dbl_sin"])
    # KAT (i): 2046 real frequencies + 2 identity pads at the FRONT: head 0, slots 0 and 1 (rope.py:499-509)
    assert torch.all(c[0, 0, :, :2] == 1) and torch.all(s[0, 0, :, :2] == 0)
    assert float(c.abs().max()) <= 1 and float(s.abs().max()) <= 1 and torch.isfinite(c).all()
    pos = torch.from_numpy(O.create_audio_position_grid(1, 68))
    c, s = O.precompute_freqs_cis(pos, 2048, 10000.0, [20], True, 32, O.LTXRopeType.SPLIT, True)
    assert np.array_equal(c.numpy(), g["audio_prod_split_dbl_cos"]) and np.array_equal(s.numpy(), g["audio_prod_split_dbl_sin"])
    x, y = torch.from_numpy(g["apply_split_x"]), torch.from_numpy(g["apply_split_y"])
    pos = torch.from_numpy(O.create_position_grid(2, 3, 4, 6))
    c, s = O.precompute_freqs_cis(pos, 512, 10000.0, [20, 2048, 2048], True, 4, O.LTXRopeType.SPLIT, True)
    out = O.apply_split_rotary_emb(x, c, s)
    assert torch.equal(out, y)
    # KAT (iii): a rotation preserves the norm of every (first-half, second-half) pair
    xh, oh = x.reshape(2, 72, 4, 2, 64), out.reshape(2, 72, 4, 2, 64)
    assert torch.allclose((xh ** 2).sum(3), (oh ** 2).sum(3), rtol=1e-5, atol=1e-6)


def test_small_ops(golden):
    g = golden("ops")
    e = O.get_timestep_embedding(torch.from_numpy(g["ts_in"]), 256, flip_sin_to_cos=True, downscale_freq_shift=0)
    assert rel_l2(e, torch.from_numpy(g["ts_emb"])) < 1e-6
    assert rel_l2(O.rms_norm(torch.from_numpy(g["rms_in"])), torch.from_numpy(g["rms_out"])) < 1e-6
    d = O.to_denoised(torch.from_numpy(g["den_noisy"]), torch.from_numpy(g["den_vel"]), torch.from_numpy(g["den_sigma"]))
    assert torch.equal(d, torch.from_numpy(g["den_out"]))


@pytest.mark.parametrize("name", sorted(CASES))
def test_model_forward_matches_reference(golden, name):
    g = golden(f"model_{name}")
    mt, L = CASES[name]
    cfg = O.small_config(mt, num_layers=L)
    tensors = O.init_params(cfg, seed=int(g["seed"]))
    assert abs(sum(float(v.double().sum()) for v in tensors.values()) - float(g["weight_checksum"])) < 1e-6
    video, audio = modalities(g)
    ov, oa = O.OracleLTXModel(cfg, tensors)(video, audio)
    if video is not None:
        assert rel_l2(ov, torch.from_numpy(g["v_out"])) < 2e-5
    if audio is not None:
        assert rel_l2(oa, torch.from_numpy(g["a_out"])) < 2e-5


def test_block_matches_reference(golden):
    g = golden("block_video")
    cfg = O.small_config(O.LTXModelType.VideoOnly, num_layers=1)
    model = O.OracleLTXModel(cfg, O.init_params(cfg, seed=int(g["seed"])))
    t = lambda k: torch.from_numpy(g[k])  # noqa: E731
    args = O.TransformerArgs(t("x_in"), t("context"), None, t("timesteps"), None, (t("cos"), t("sin")), None, None, None, True)
    out, _ = model.block(0, args, None)
    assert rel_l2(out.x, t("x_out")) < 2e-5


def test_zero_gates_reduce_block_to_text_cross_attention():
    # KAT (ii): zero tables + zero timestep modulation => attn1 and ff are gated off; x + attn2(rms(x), ctx) remains
    cfg = O.small_config(O.LTXModelType.VideoOnly, num_layers=1)
    tensors = O.init_params(cfg, seed=3, table_std=0.0)
    model = O.OracleLTXModel(cfg, tensors)
    g = torch.Generator().manual_seed(4)
    x, ctx = torch.randn(1, 24, 512, generator=g), torch.randn(1, 8, 512, generator=g)
    pos = torch.from_numpy(O.create_position_grid(1, 2, 3, 4))
    pe = O.precompute_freqs_cis(pos, 512, 10000.0, [20, 2048, 2048], True, 4, O.LTXRopeType.SPLIT, True)
    args = O.TransformerArgs(x, ctx, None, torch.zeros(1, 24, 6 * 512), None, pe, None, None, None, True)
    out, _ = model.block(0, args, None)
    p = O.Params(tensors).sub("transformer_blocks.0")
    expect = x + O.attention(p.sub("attn2"), O.rms_norm(x), 4, cfg.rope_type, cfg.norm_eps, context=ctx)
    assert rel_l2(out.x, expect) < 1e-6


def test_error_behaviour():
    cfg = O.small_config(O.LTXModelType.VideoOnly, num_layers=1)
    model = O.OracleLTXModel(cfg, O.init_params(cfg, seed=1))
    m = O.Modality(torch.zeros(1, 4, 128), torch.zeros(1, 4), torch.zeros(1, 1, 4, 2), torch.zeros(1, 2, 256))
    with pytest.raises(ValueError):  # ltx.py:468-469
        model(video=None, audio=m)
    with pytest.raises(AssertionError):  # rope.py:228 — axes vs max_pos
        O.precompute_freqs_cis(torch.zeros(1, 2, 4, 2), 512, 10000.0, [20, 2048, 2048], True, 4, O.LTXRopeType.SPLIT, True)

"""ORACLE / test infrastructure only — generates tests/golden/*.npz.

Runs the reference's OWN Python sources (/root/reference/mlx_video/models/ltx/*.py and the sampler
helpers of mlx_video/generate.py), imported unmodified over ``oracle/mlx_shim`` (ref_loader.py), on
seeded inputs and the oracle's random-init weights, and stores inputs + outputs as small fixtures.
It also asserts, while generating, that ``oracle/ltx_oracle.py`` reproduces every reference output
(bit-exact for the integer/position work, <= 2e-5 relative for fp32 arithmetic whose summation
order differs).  /root/reference does not exist on the GPU box, so this script only runs in the
build container; the committed fixtures are what travels.

    python oracle/make_golden.py            # regenerate + self-check
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))

import ltx_oracle as O  # noqa: E402
import ref_loader  # noqa: E402

GOLDEN = HERE.parent / "tests" / "golden"


def rel(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def ref_config(R, cfg: O.OracleConfig):
    """Field-for-field copy of an OracleConfig into the reference's LTXModelConfig (config.py:93-135)."""
    C = R.config
    mt = {O.LTXModelType.VideoOnly: C.LTXModelType.VideoOnly, O.LTXModelType.AudioVideo: C.LTXModelType.AudioVideo,
          O.LTXModelType.AudioOnly: C.LTXModelType.AudioOnly}[cfg.model_type]
    rt = {O.LTXRopeType.SPLIT: C.LTXRopeType.SPLIT, O.LTXRopeType.INTERLEAVED: C.LTXRopeType.INTERLEAVED}[cfg.rope_type]
    return C.LTXModelConfig(
        model_type=mt, num_attention_heads=cfg.num_attention_heads, attention_head_dim=cfg.attention_head_dim,
        in_channels=cfg.in_channels, out_channels=cfg.out_channels, num_layers=cfg.num_layers,
        cross_attention_dim=cfg.cross_attention_dim, caption_channels=cfg.caption_channels,
        audio_num_attention_heads=cfg.audio_num_attention_heads, audio_attention_head_dim=cfg.audio_attention_head_dim,
        audio_in_channels=cfg.audio_in_channels, audio_out_channels=cfg.audio_out_channels,
        audio_cross_attention_dim=cfg.audio_cross_attention_dim, audio_caption_channels=cfg.audio_caption_channels,
        positional_embedding_theta=cfg.positional_embedding_theta,
        positional_embedding_max_pos=list(cfg.positional_embedding_max_pos),
        audio_positional_embedding_max_pos=list(cfg.audio_positional_embedding_max_pos),
        use_middle_indices_grid=cfg.use_middle_indices_grid, rope_type=rt,
        double_precision_rope=cfg.double_precision_rope, timestep_scale_multiplier=cfg.timestep_scale_multiplier,
        av_ca_timestep_scale_multiplier=cfg.av_ca_timestep_scale_multiplier, norm_eps=cfg.norm_eps)


def build_reference_model(R, cfg: O.OracleConfig, tensors):
    model = R.ltx.LTXModel(ref_config(R, cfg))
    for name, value in tensors.items():
        ref_loader.set_param(model, name, R.mx.array(value.clone()))
    return model


def ref_modality(R, m: O.Modality):
    a = R.mx.array
    pe = None if m.positional_embeddings is None else (a(m.positional_embeddings[0]), a(m.positional_embeddings[1]))
    return R.transformer.Modality(latent=a(m.latent), timesteps=a(m.timesteps), positions=a(m.positions),
                                  context=a(m.context), enabled=m.enabled,
                                  context_mask=None if m.context_mask is None else a(m.context_mask),
                                  positional_embeddings=pe)


def make_inputs(cfg: O.OracleConfig, seed: int, B: int, grid, Tc: int, sigma: float, Ta: int = 0, per_frame_mask=False,
                with_mask=False):
    """Seeded synthetic modalities (SURVEY.md §8c random-init scheme)."""
    g = torch.Generator().manual_seed(seed)
    F_, H_, W_ = grid
    T = F_ * H_ * W_
    video = audio = None
    if cfg.model_type.is_video_enabled():
        lat = torch.randn(B, T, cfg.in_channels, generator=g)
        ctx = torch.randn(B, Tc, cfg.caption_channels, generator=g)
        ts = torch.full((B, T), sigma)
        if per_frame_mask:  # I2V: frame 0 is clean (mask 0) -> timestep 0 on its tokens (conditioning/latent.py:166)
            ts[:, : H_ * W_] = 0.0
        pos = torch.from_numpy(O.create_position_grid(B, F_, H_, W_))
        cmask = None
        if with_mask:
            cmask = torch.ones(B, Tc, dtype=torch.int32)
            cmask[:, Tc - Tc // 4:] = 0
        video = O.Modality(lat, ts, pos, ctx, True, cmask)
    if cfg.model_type.is_audio_enabled():
        lat = torch.randn(B, Ta, cfg.audio_in_channels, generator=g)
        ctx = torch.randn(B, Tc, cfg.audio_caption_channels, generator=g)
        ts = torch.full((B, Ta), sigma)
        pos = torch.from_numpy(O.create_audio_position_grid(B, Ta))
        audio = O.Modality(lat, ts, pos, ctx, True, None)
    return video, audio


def np_(t):
    return None if t is None else t.detach().cpu().numpy()


def main() -> int:
    torch.set_num_threads(8)
    R = ref_loader.load()
    GOLDEN.mkdir(parents=True, exist_ok=True)
    worst = {}

    # ---------------------------------------------------------------- grids / schedulers (bit-exact)
    grids = {}
    for name, args in {"g_1_5_16_16": (1, 5, 16, 16), "g_2_3_4_6": (2, 3, 4, 6), "g_1_9_24_24": (1, 9, 24, 24)}.items():
        ref = np.asarray(R.generate.create_position_grid(*args))
        mine = O.create_position_grid(*args)
        assert ref.dtype == np.float32 and np.array_equal(ref, mine), name
        grids[name] = ref if ref.size < 20000 else ref[:, :, ::7, :]  # big grid: every 7th token
    for name, args in {"a_1_68": (1, 68), "a_2_21": (2, 21)}.items():
        ref = np.asarray(R.generate.create_audio_position_grid(*args))
        assert np.array_equal(ref, O.create_audio_position_grid(*args)), name
        grids[name] = ref
    for steps, ntok in [(40, 5184), (8, 1280), (30, None), (4, 320)]:
        ref = np.asarray(R.generate.ltx2_scheduler(steps, ntok))
        assert np.array_equal(ref.astype(np.float32), O.ltx2_scheduler(steps, ntok)), (steps, ntok)
        grids[f"sched_{steps}_{ntok}"] = ref.astype(np.float32)
    grids["STAGE_1_SIGMAS"] = np.asarray(R.generate.STAGE_1_SIGMAS, dtype=np.float64)
    grids["STAGE_2_SIGMAS"] = np.asarray(R.generate.STAGE_2_SIGMAS, dtype=np.float64)
    assert list(grids["STAGE_1_SIGMAS"]) == O.STAGE_1_SIGMAS and list(grids["STAGE_2_SIGMAS"]) == O.STAGE_2_SIGMAS
    assert R.generate.compute_audio_frames(65, 24.0) == O.compute_audio_frames(65, 24.0) == 68
    np.savez_compressed(GOLDEN / "grids.npz", **grids)

    # ---------------------------------------------------------------- RoPE tables
    rope = {}
    RT = R.config.LTXRopeType
    for name, (grid_args, dim, heads, max_pos, audio) in {
        "video_prod": ((1, 5, 16, 16), 4096, 32, [20, 2048, 2048], False),
        "video_small": ((2, 3, 4, 6), 512, 4, [20, 2048, 2048], False),
        "audio_prod": ((1, 68), 2048, 32, [20], True),
        "cross_small": ((1, 21), 256, 4, [20], True),
    }.items():
        pos = O.create_audio_position_grid(*grid_args) if audio else O.create_position_grid(*grid_args)
        for rt_name, rt_ref, rt_mine in [("split", RT.SPLIT, O.LTXRopeType.SPLIT), ("inter", RT.INTERLEAVED, O.LTXRopeType.INTERLEAVED)]:
            for dbl in (True, False):
                c_ref, s_ref = R.rope.precompute_freqs_cis(R.mx.array(pos), dim=dim, theta=10000.0, max_pos=max_pos,
                                                           use_middle_indices_grid=True, num_attention_heads=heads,
                                                           rope_type=rt_ref, double_precision=dbl)
                c_me, s_me = O.precompute_freqs_cis(torch.from_numpy(pos), dim, 10000.0, max_pos, True, heads, rt_mine, dbl)
                key = f"{name}_{rt_name}_{'dbl' if dbl else 'std'}"
                dc, ds = (c_ref._t - c_me).abs().max().item(), (s_ref._t - s_me).abs().max().item()
                worst["rope_" + key] = max(dc, ds)
                # the "double precision" path is bit-reproducible; the plain path goes through a different
                # (but mathematically identical) op order in the reference
                assert max(dc, ds) <= (0.0 if dbl else 2e-3), (key, dc, ds)
                if dbl and rt_name == "split":
                    c, s = np_(c_ref._t), np_(s_ref._t)
                    if c.size > 200000:  # production table: keep heads {0,1,17,31}, every 5th token
                        c, s = c[:, [0, 1, 17, 31], ::5], s[:, [0, 1, 17, 31], ::5]
                    rope[key + "_cos"], rope[key + "_sin"] = c, s
    # apply_split_rotary_emb / interleaved on a (B,T,H*dh) tensor
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 72, 512, generator=g)
    pos = O.create_position_grid(2, 3, 4, 6)
    c, s = O.precompute_freqs_cis(torch.from_numpy(pos), 512, 10000.0, [20, 2048, 2048], True, 4, O.LTXRopeType.SPLIT, True)
    y_ref = R.rope.apply_split_rotary_emb(R.mx.array(x), R.mx.array(c), R.mx.array(s))._t
    assert torch.equal(y_ref, O.apply_split_rotary_emb(x, c, s))
    rope["apply_split_x"], rope["apply_split_y"] = np_(x), np_(y_ref)
    ci, si = O.precompute_freqs_cis(torch.from_numpy(pos), 512, 10000.0, [20, 2048, 2048], True, 4, O.LTXRopeType.INTERLEAVED, True)
    yi_ref = R.rope.apply_interleaved_rotary_emb(R.mx.array(x), R.mx.array(ci), R.mx.array(si))._t
    assert rel(O.apply_interleaved_rotary_emb(x, ci, si), yi_ref) < 1e-6
    np.savez_compressed(GOLDEN / "rope.npz", **rope)

    # ---------------------------------------------------------------- small ops
    ops = {}
    t = torch.tensor([0.0, 1.0, 50.0, 421.875, 725.0, 993.75, 1000.0])
    e_ref = R.utils.get_timestep_embedding(R.mx.array(t), 256, flip_sin_to_cos=True, downscale_freq_shift=0)._t
    e_me = O.get_timestep_embedding(t, 256, flip_sin_to_cos=True, downscale_freq_shift=0)
    assert rel(e_me, e_ref) < 1e-6
    ops["ts_in"], ops["ts_emb"] = np_(t), np_(e_ref)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(3, 7, 512, generator=g) * 3
    r_ref = R.utils.rms_norm(R.mx.array(x))._t
    assert rel(O.rms_norm(x), r_ref) < 1e-6
    ops["rms_in"], ops["rms_out"] = np_(x), np_(r_ref)
    noisy, vel = torch.randn(2, 9, 128, generator=g), torch.randn(2, 9, 128, generator=g)
    sig = torch.rand(2, 9, generator=g)
    d_ref = R.utils.to_denoised(R.mx.array(noisy), R.mx.array(vel), R.mx.array(sig))._t
    assert torch.equal(d_ref, O.to_denoised(noisy, vel, sig))
    ops["den_noisy"], ops["den_vel"], ops["den_sigma"], ops["den_out"] = np_(noisy), np_(vel), np_(sig), np_(d_ref)
    np.savez_compressed(GOLDEN / "ops.npz", **ops)

    # ---------------------------------------------------------------- model forwards (reference code over the shim)
    cases = {
        # name: (config, B, grid, Tc, sigma, Ta, per_frame_mask, with_mask)
        "video_L2": (O.small_config(O.LTXModelType.VideoOnly, num_layers=2), 1, (3, 4, 6), 24, 0.725, 0, False, False),
        "video_L2_b2_i2v": (O.small_config(O.LTXModelType.VideoOnly, num_layers=2), 2, (2, 4, 5), 16, 0.421875, 0, True, False),
        "video_L1_mask": (O.small_config(O.LTXModelType.VideoOnly, num_layers=1), 2, (2, 3, 4), 16, 1.0, 0, False, True),
        "av_L2": (O.small_config(O.LTXModelType.AudioVideo, num_layers=2), 1, (3, 4, 6), 24, 0.725, 21, False, False),
        "audio_L1": (O.small_config(O.LTXModelType.AudioOnly, num_layers=1), 1, (1, 1, 1), 16, 0.05, 21, False, False),
    }
    for name, (cfg, B, grid, Tc, sigma, Ta, pfm, wm) in cases.items():
        seed = abs(hash(name)) % 1000 if False else sum(map(ord, name))
        tensors = O.init_params(cfg, seed=seed)
        video, audio = make_inputs(cfg, seed + 1, B, grid, Tc, sigma, Ta, pfm, wm)
        ref_model = build_reference_model(R, cfg, tensors)
        rv, ra = ref_model(video=None if video is None else ref_modality(R, video),
                           audio=None if audio is None else ref_modality(R, audio))
        ov, oa = O.OracleLTXModel(cfg, tensors)(video, audio)
        out = {"seed": np.int64(seed)}
        if video is not None:
            worst[name + "_video"] = rel(ov, rv._t)
            assert worst[name + "_video"] < 2e-5, (name, worst[name + "_video"])
            out.update(v_latent=np_(video.latent), v_timesteps=np_(video.timesteps), v_positions=np_(video.positions),
                       v_context=np_(video.context), v_out=np_(rv._t))
            if video.context_mask is not None:
                out["v_context_mask"] = np_(video.context_mask)
        if audio is not None:
            worst[name + "_audio"] = rel(oa, ra._t)
            assert worst[name + "_audio"] < 2e-5, (name, worst[name + "_audio"])
            out.update(a_latent=np_(audio.latent), a_timesteps=np_(audio.timesteps), a_positions=np_(audio.positions),
                       a_context=np_(audio.context), a_out=np_(ra._t))
        # weight fingerprint so a drift of the RNG / init scheme is caught before a parity failure is blamed on kernels
        out["weight_checksum"] = np.float64(sum(float(v.double().sum()) for v in tensors.values()))
        np.savez_compressed(GOLDEN / f"model_{name}.npz", **out)

    # single block, reference BasicAVTransformerBlock called directly (transformer.py:221-361)
    cfg = O.small_config(O.LTXModelType.VideoOnly, num_layers=1)
    tensors = O.init_params(cfg, seed=77)
    video, _ = make_inputs(cfg, 78, 1, (2, 4, 4), 16, 0.725)
    oracle_model = O.OracleLTXModel(cfg, tensors)
    va, _ = oracle_model.prepare(video, None)
    ref_model = build_reference_model(R, cfg, tensors)
    rva = ref_model.video_args_preprocessor.prepare(ref_modality(R, video))
    assert rel(va.x, rva.x._t) < 1e-6 and rel(va.timesteps, rva.timesteps._t) < 1e-5 and rel(va.context, rva.context._t) < 1e-5
    rb, _ = ref_model.transformer_blocks[0](video=rva, audio=None)
    ob, _ = oracle_model.block(0, va, None)
    worst["block_video"] = rel(ob.x, rb.x._t)
    assert worst["block_video"] < 2e-5
    np.savez_compressed(GOLDEN / "block_video.npz", x_in=np_(rva.x._t), timesteps=np_(rva.timesteps._t),
                        context=np_(rva.context._t), cos=np_(rva.positional_embeddings[0]._t),
                        sin=np_(rva.positional_embeddings[1]._t), x_out=np_(rb.x._t), seed=np.int64(77))

    # reference error behaviour (ltx.py:466-469)
    try:
        ref_model(video=None, audio=ref_modality(R, video))
        raise AssertionError("reference accepted audio on a video-only model")
    except ValueError:
        pass

    for k, v in sorted(worst.items()):
        print(f"{k:40s} oracle-vs-reference {v:.3e}")
    total = sum(p.stat().st_size for p in GOLDEN.glob("*.npz"))
    print(f"wrote {len(list(GOLDEN.glob('*.npz')))} fixtures, {total / 1e6:.2f} MB, to {GOLDEN}")
    return 0


if __name__ == "__main__":
    sys.exit(main())

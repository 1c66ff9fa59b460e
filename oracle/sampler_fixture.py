"""ORACLE / test infrastructure only — seeded inputs of the denoise-loop parity cases (tests/golden/sampler.npz).
Shared by oracle/make_golden_sampler.py (which runs the reference's own loops on them) and tests/ (oracle and product
loops on the same inputs).  Everything is regenerated from the seeds; only the reference's outputs are committed."""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch

import ltx_oracle as O

STAGE_2_SIGMAS = [0.909375, 0.725, 0.421875, 0.0]  # generate.py:353
GRID = (2, 4, 6)       # F, H, W -> 48 video tokens
TC, TA = 16, 9         # text tokens, audio latent frames

# case -> model type, layers, weight seed, loop, loop arguments
CASES: Dict[str, dict] = {
    "distilled":       dict(mt="video", L=2, seed=31, loop="distilled", sigmas=STAGE_2_SIGMAS),
    "distilled_i2v":   dict(mt="video", L=2, seed=31, loop="distilled", sigmas=STAGE_2_SIGMAS, state=True),
    "dev":             dict(mt="video", L=2, seed=31, loop="dev", steps=3, cfg_scale=4.5),
    "dev_i2v":         dict(mt="video", L=2, seed=31, loop="dev", steps=2, cfg_scale=3.0, state=True),
    "dev_nocfg":       dict(mt="video", L=2, seed=31, loop="dev", steps=2, cfg_scale=1.0),
    "distilled_av":    dict(mt="av", L=1, seed=32, loop="distilled", sigmas=STAGE_2_SIGMAS, audio=True),
    "dev_av":          dict(mt="av", L=1, seed=32, loop="dev_av", steps=2, cfg_scale=3.0),
    "dev_av_i2v":      dict(mt="av", L=1, seed=32, loop="dev_av", steps=2, cfg_scale=4.0, state=True),
    "audio_only":      dict(mt="audio", L=1, seed=33, loop="audio_only", sigmas=STAGE_2_SIGMAS),
}
_MT = {"video": O.LTXModelType.VideoOnly, "av": O.LTXModelType.AudioVideo, "audio": O.LTXModelType.AudioOnly}


def config(case: str) -> O.OracleConfig:
    c = CASES[case]
    return O.small_config(_MT[c["mt"]], num_layers=c["L"])


def weights(case: str, bf16_exact: bool = True) -> Dict[str, torch.Tensor]:
    """Seeded weights; linear weights rounded to bf16 values so the GPU model (bf16 storage) holds the SAME numbers."""
    t = O.init_params(config(case), seed=CASES[case]["seed"])
    if bf16_exact:
        t = {k: (v.to(torch.bfloat16).float() if k.endswith(".weight") and not k.endswith("_norm.weight") else v) for k, v in t.items()}
    return t


def sigmas(case: str) -> np.ndarray:
    c = CASES[case]
    if "sigmas" in c:
        return np.asarray(c["sigmas"], np.float32)
    F_, H_, W_ = GRID
    return O.ltx2_scheduler(c["steps"], F_ * H_ * W_)


def inputs(case: str) -> Dict[str, torch.Tensor]:
    """Latents in the reference layouts (video (B,128,F,H,W), audio (B,8,Ta,16)), position grids, text embeddings,
    and — for the i2v cases — a LatentState whose first frame is conditioning (mask 0.25: partly denoised)."""
    c = CASES[case]
    cfg = config(case)
    g = torch.Generator().manual_seed(c["seed"] + 1000)
    F_, H_, W_ = GRID
    out: Dict[str, torch.Tensor] = {}
    if cfg.model_type.is_video_enabled():
        out["latents"] = torch.randn(1, 128, F_, H_, W_, generator=g)
        out["positions"] = torch.from_numpy(O.create_position_grid(1, F_, H_, W_))
        out["ctx_pos"] = torch.randn(1, TC, cfg.caption_channels, generator=g)
        out["ctx_neg"] = torch.randn(1, TC, cfg.caption_channels, generator=g)
        if c.get("state"):
            out["clean"] = torch.randn(1, 128, F_, H_, W_, generator=g)
            mask = torch.ones(1, 1, F_, 1, 1)
            mask[:, :, 0] = 0.25
            out["mask"] = mask
    if cfg.model_type.is_audio_enabled():
        out["audio_latents"] = torch.randn(1, 8, TA, 16, generator=g)
        out["audio_positions"] = torch.from_numpy(O.create_audio_position_grid(1, TA))
        out["actx_pos"] = torch.randn(1, TC, cfg.audio_caption_channels, generator=g)
        out["actx_neg"] = torch.randn(1, TC, cfg.audio_caption_channels, generator=g)
    return out


def run_oracle(case: str, cfg_batch: bool = False):
    """The oracle's restatement of the case's loop -> (video latents or None, audio latents or None)."""
    c, x = CASES[case], inputs(case)
    model = O.OracleLTXModel(config(case), weights(case))
    sig = sigmas(case)
    state = O.LatentState(x["latents"], x["clean"], x["mask"]) if c.get("state") else None
    if c["loop"] == "distilled":
        if c.get("audio"):
            return O.denoise_distilled(x["latents"], x["positions"], x["ctx_pos"], model, sig, state=state,
                                       audio_latents=x["audio_latents"], audio_positions=x["audio_positions"],
                                       audio_embeddings=x["actx_pos"])
        return O.denoise_distilled(x["latents"], x["positions"], x["ctx_pos"], model, sig, state=state)
    if c["loop"] == "dev":
        return O.denoise_dev(x["latents"], x["positions"], x["ctx_pos"], x["ctx_neg"], model, sig, cfg_scale=c["cfg_scale"],
                             state=state, cfg_batch=cfg_batch), None
    if c["loop"] == "dev_av":
        return O.denoise_dev_av(x["latents"], x["audio_latents"], x["positions"], x["audio_positions"], x["ctx_pos"], x["ctx_neg"],
                                x["actx_pos"], x["actx_neg"], model, sig, cfg_scale=c["cfg_scale"], video_state=state,
                                cfg_batch=cfg_batch)
    if c["loop"] == "audio_only":
        return None, O.denoise_audio_only(x["audio_latents"], x["audio_positions"], x["actx_pos"], model, sig)
    raise KeyError(c["loop"])

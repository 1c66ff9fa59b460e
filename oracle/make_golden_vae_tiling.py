"""ORACLE / test infrastructure only — generates tests/golden/vae_tiling.npz by running the reference's OWN
``decode_with_tiling`` (mlx_video/models/ltx/video_vae/tiling.py:299-520), ``split_in_*``, ``map_*_slice`` and
``compute_trapezoidal_mask_1d`` unmodified over oracle/mlx_shim.  The decoder is replaced by ``fake_decoder`` below — a cheap
function of the tile's latents AND of tile-local coordinates, so overlapping tiles disagree and the blend weights show —
because what is pinned here is the tiling / blending logic, not the network (tests/golden/vae_decoder.npz pins that).

    python oracle/make_golden_vae_tiling.py
"""
from __future__ import annotations

import importlib
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))

import ref_loader  # noqa: E402

GOLDEN = HERE.parent / "tests" / "golden"
CASES = {
    # name: latent shape, spatial (tile px, overlap px) or None, temporal (tile frames, overlap frames) or None
    "both": dict(shape=(1, 128, 5, 5, 6), spatial=(64, 32), temporal=(16, 8)),
    "spatial": dict(shape=(2, 128, 2, 7, 4), spatial=(96, 32), temporal=None),
    "temporal": dict(shape=(1, 128, 9, 2, 2), spatial=None, temporal=(24, 8)),
}


SAMPLE = (3, 5)  # stored rows / columns of every frame


def case_latents(name: str) -> torch.Tensor:
    g = torch.Generator().manual_seed(77 + len(name))
    return torch.randn(*CASES[name]["shape"], generator=g)


def fake_decoder_torch(tile: torch.Tensor) -> torch.Tensor:
    """(B, C, f, h, w) -> (B, 3, 1 + 8(f-1), 32h, 32w): nearest-neighbour expansion of the first three channels (output frame k
    shows latent frame ceil(k / 8)) plus ramps in the TILE-LOCAL output coordinates."""
    B, _, f, h, w = tile.shape
    F_ = 1 + 8 * (f - 1)
    fi = torch.div(torch.arange(F_) + 7, 8, rounding_mode="floor")
    up = tile[:, :3][:, :, fi].repeat_interleave(32, dim=3).repeat_interleave(32, dim=4)
    t = torch.arange(F_, dtype=torch.float32).reshape(1, 1, F_, 1, 1)
    y = torch.arange(32 * h, dtype=torch.float32).reshape(1, 1, 1, -1, 1)
    x = torch.arange(32 * w, dtype=torch.float32).reshape(1, 1, 1, 1, -1)
    return up + 0.05 * t + 0.003 * y - 0.002 * x


def main() -> int:
    R = ref_loader.load()
    ref_loader._stub_package("mlx_video.models.ltx.video_vae", ref_loader.REFERENCE_ROOT / "mlx_video" / "models" / "ltx" / "video_vae")
    til = importlib.import_module("mlx_video.models.ltx.video_vae.tiling")
    a = R.mx.array

    def decoder_fn(tile, causal=False, timestep=None, debug=False, chunked_conv=False):
        return a(fake_decoder_torch(tile._t))

    out = {}
    for name, c in CASES.items():
        cfg = til.TilingConfig(
            spatial_config=None if c["spatial"] is None else til.SpatialTilingConfig(*c["spatial"]),
            temporal_config=None if c["temporal"] is None else til.TemporalTilingConfig(*c["temporal"]))
        emitted = []
        ref = til.decode_with_tiling(decoder_fn, a(case_latents(name)), cfg, on_frames_ready=lambda fr, start: emitted.append((int(start), tuple(fr.shape))))
        out[f"{name}/output"] = ref._t[:, :, :, ::SAMPLE[0], ::SAMPLE[1]].contiguous().numpy()  # strided sample keeps the fixture small
        out[f"{name}/emitted"] = np.array([[s, sh[2]] for s, sh in emitted], dtype=np.int64)
        print(name, tuple(ref.shape), "emitted", emitted)
    # the mask / split helpers on a sweep of arguments
    masks, splits = [], []
    for length, rl, rr, z in [(64, 0, 32, False), (64, 32, 0, False), (17, 9, 8, True), (17, 0, 8, True), (8, 8, 8, False), (5, 9, 2, True)]:
        m = til.compute_trapezoidal_mask_1d(length, rl, rr, z)._t.numpy()
        masks.append(np.concatenate([[length, rl, rr, int(z)], m]).astype(np.float32))
    out["masks"] = np.array(masks, dtype=object)
    for fn in ("split_in_spatial", "split_in_temporal"):
        for size, ov, dim in [(2, 1, 5), (3, 1, 7), (2, 1, 9), (4, 2, 4), (8, 3, 30), (3, 1, 2)]:
            iv = getattr(til, fn)(size, ov, dim)
            splits.append((fn, size, ov, dim, list(iv.starts), list(iv.ends), list(iv.left_ramps), list(iv.right_ramps)))
    out["splits"] = np.array(splits, dtype=object)
    np.savez_compressed(GOLDEN / "vae_tiling.npz", **out)
    print(f"wrote {GOLDEN / 'vae_tiling.npz'}")
    return 0


if __name__ == "__main__":
    sys.exit(main())

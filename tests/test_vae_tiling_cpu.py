"""CPU: the host-side tiling logic of the VAE decoder (mlx-video_b200/vae_decoder.py: trapezoid masks, interval splits,
TilingConfig presets) against what the reference's OWN tiling.py produced over the shim (tests/golden/vae_tiling.npz,
oracle/make_golden_vae_tiling.py).  Integer / index work: exact; masks: fp32-exact."""
import numpy as np
import pytest

import mlx_video_b200  # noqa: F401
from mlx_video_b200 import vae_decoder as VD


def test_masks_and_splits_match_reference(golden):
    import pathlib
    g = np.load(pathlib.Path(__file__).parent / "golden" / "vae_tiling.npz", allow_pickle=True)
    for row in g["masks"]:
        length, rl, rr, z = (int(v) for v in row[:4])
        assert np.array_equal(VD.compute_trapezoidal_mask_1d(length, rl, rr, bool(z)), row[4:].astype(np.float32)), (length, rl, rr, z)
    for fn, size, ov, dim, starts, ends, lr, rr in g["splits"]:
        iv = getattr(VD, fn)(int(size), int(ov), int(dim))
        assert (iv.starts, iv.ends, iv.left_ramps, iv.right_ramps) == (list(starts), list(ends), list(lr), list(rr)), (fn, size, ov, dim)


def test_tiling_config_validation_and_presets():
    for bad in (dict(tile_size_in_pixels=32), dict(tile_size_in_pixels=100), dict(tile_size_in_pixels=64, tile_overlap_in_pixels=16),
                dict(tile_size_in_pixels=64, tile_overlap_in_pixels=64)):
        with pytest.raises(ValueError):
            VD.SpatialTilingConfig(**bad)
    for bad in (dict(tile_size_in_frames=8), dict(tile_size_in_frames=20), dict(tile_size_in_frames=16, tile_overlap_in_frames=4),
                dict(tile_size_in_frames=16, tile_overlap_in_frames=16)):
        with pytest.raises(ValueError):
            VD.TemporalTilingConfig(**bad)
    d = VD.TilingConfig.default()
    assert (d.spatial_config.tile_size_in_pixels, d.spatial_config.tile_overlap_in_pixels) == (512, 64)
    assert (d.temporal_config.tile_size_in_frames, d.temporal_config.tile_overlap_in_frames) == (64, 24)
    assert VD.TilingConfig.auto(512, 512, 33) is None                      # tiling.py:175-176
    a = VD.TilingConfig.auto(768, 768, 65)
    assert a.spatial_config.tile_size_in_pixels == 384 and a.temporal_config is None
    assert VD.TilingConfig.auto(704, 1280, 121) == VD.TilingConfig.aggressive()  # > 768x1024 pixels and > 100 frames (tiling.py:185-186)
    a = VD.TilingConfig.auto(704, 1280, 97)
    assert a.spatial_config.tile_size_in_pixels == 384 and (a.temporal_config.tile_size_in_frames, a.temporal_config.tile_overlap_in_frames) == (64, 24)
    a = VD.TilingConfig.auto(512, 1024, 121)
    assert a.spatial_config.tile_size_in_pixels == 512 and (a.temporal_config.tile_size_in_frames, a.temporal_config.tile_overlap_in_frames) == (48, 16)
    assert VD.TilingConfig.auto(1088, 1920, 257) == VD.TilingConfig.aggressive()
    t, m = VD.map_temporal_slice(1, 3, 2, 1, 8)
    assert (t.start, t.stop) == (8, 17) and m.shape == (9,) and m[0] == 0.0
    s, m = VD.map_spatial_slice(1, 3, 1, 0, 32)
    assert (s.start, s.stop) == (32, 96) and m[-1] == 1.0 and 0 < m[0] < m[31] < 1.0

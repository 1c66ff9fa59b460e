"""ORACLE / test infrastructure only — generates tests/golden/vae_decoder.npz by running the reference's OWN
``LTX2VideoDecoder`` (mlx_video/models/ltx/video_vae/decoder.py, full width: 1024 / 512 / 256 / 128 channels, 545 M seeded
parameters) unmodified over oracle/mlx_shim on two tiny latents, and asserts that oracle/vae_decoder_oracle.py reproduces
both outputs (<= 2e-5 relative).  Groundwork for SURVEY.md §8f row N4 (second half): there is no product path yet.

    python oracle/make_golden_vae.py    # needs /root/reference (this container) and ~8 GB of RAM

Cases: ``plain`` — non-causal, noise off, default decode timestep; ``causal_noise`` — causal temporal padding, the
decoder's own noise injection (scale 0.025; the N(0,1) draw is reproduced from the same torch seed), explicit timestep.
Stored: the full first output, a strided sample of the second.
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
import vae_decoder_oracle as V  # noqa: E402
from make_golden import rel  # noqa: E402

GOLDEN = HERE.parent / "tests" / "golden"
SEED = 51
CASES = {
    "plain": dict(shape=(1, 128, 2, 2, 2), causal=False, noise_scale=0.0, timestep=None, sample=(1, 1, 1)),
    "causal_noise": dict(shape=(1, 128, 3, 2, 3), causal=True, noise_scale=0.025, timestep=0.1, sample=(2, 3, 4)),
}


def case_inputs(name: str):
    c = CASES[name]
    g = torch.Generator().manual_seed(SEED + len(name))
    return torch.randn(*c["shape"], generator=g)


def main() -> int:
    torch.set_num_threads(8)
    R = ref_loader.load()
    ref_loader._stub_package("mlx_video.models.ltx.video_vae", ref_loader.REFERENCE_ROOT / "mlx_video" / "models" / "ltx" / "video_vae")
    dec = importlib.import_module("mlx_video.models.ltx.video_vae.decoder")
    a = R.mx.array
    params = V.init_decoder_params(SEED)
    model = dec.LTX2VideoDecoder()
    from mlx.utils import tree_flatten

    names = set(dict(tree_flatten(model.parameters())))
    assert names == set(params), sorted(names ^ set(params))[:8]
    for name, value in params.items():
        ref_loader.set_param(model, name, a(value))
    out = {}
    for name, c in CASES.items():
        x = case_inputs(name)
        model.decode_noise_scale = c["noise_scale"]
        ts = None if c["timestep"] is None else a(torch.full((x.shape[0],), c["timestep"]))
        torch.manual_seed(1234)  # the shim's mx.random.normal draws torch.randn(shape) from the global generator
        ref = model(a(x), causal=c["causal"], timestep=ts)._t
        torch.manual_seed(1234)
        noise = torch.randn(x.shape)
        mine = V.decode(params, x, causal=c["causal"], timestep=None if ts is None else ts._t, noise=noise, noise_scale=c["noise_scale"])
        B, _, F_, H, W = x.shape
        assert tuple(ref.shape) == (B, 3, 8 * (F_ - 1) + 1, 32 * H, 32 * W), ref.shape
        r = rel(mine, ref)
        assert r <= 2e-5, (name, r)
        sf, sh, sw = c["sample"]
        out[f"{name}/output"] = ref[:, :, ::sf, ::sh, ::sw].contiguous().numpy()
        print(f"{name:13s} {tuple(ref.shape)} oracle vs reference {r:.2e}, |out| {float(ref.norm()):.1f}, stored {out[f'{name}/output'].shape}")
    np.savez_compressed(GOLDEN / "vae_decoder.npz", **out)
    print(f"wrote {GOLDEN / 'vae_decoder.npz'}")
    return 0


if __name__ == "__main__":
    sys.exit(main())

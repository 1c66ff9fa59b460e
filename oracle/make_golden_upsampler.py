"""ORACLE / test infrastructure only — generates tests/golden/upsampler.npz by running the reference's OWN
``mlx_video/models/ltx/upsampler.py`` (``LatentUpsampler``, ``upsample_latents``, and for one case ``load_upsampler`` on a
safetensors file in the upstream conv layouts) unmodified over oracle/mlx_shim on the seeded cases of
oracle/upsampler_fixture.py, and asserts that oracle/upsampler_oracle.py reproduces every output (<= 2e-5 relative).

    python oracle/make_golden_upsampler.py    # needs /root/reference (this container); the fixture is committed
"""
from __future__ import annotations

import importlib
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))

import ref_loader  # noqa: E402
import upsampler_fixture as UF  # noqa: E402
import upsampler_oracle as U  # noqa: E402,F401
from make_golden import rel  # noqa: E402

GOLDEN = HERE.parent / "tests" / "golden"


def main() -> int:
    torch.set_num_threads(8)
    R = ref_loader.load()
    up = importlib.import_module("mlx_video.models.ltx.upsampler")
    from safetensors.torch import save_file

    a = R.mx.array
    out = {}
    for case, c in UF.CASES.items():
        latent, mean, std = UF.inputs(case)
        if case == "loaded":
            with tempfile.TemporaryDirectory() as td:
                path = Path(td) / "upsampler.safetensors"
                save_file(UF.upstream_state(case), str(path))
                model = up.load_upsampler(str(path))
            assert model.mid_channels == c["mid"] and len(model.res_blocks) == 4
        else:
            model = up.LatentUpsampler(in_channels=128, mid_channels=c["mid"], num_blocks_per_stage=c["blocks"])
            for name, value in UF.params(case).items():
                ref_loader.set_param(model, name, a(value.clone()))
        ref = up.upsample_latents(a(latent), model, a(mean), a(std))._t
        mine = UF.run_oracle(case)
        b, ch, f, h, w = c["shape"]
        assert tuple(ref.shape) == (b, ch, f, 2 * h, 2 * w) and ref.dtype == torch.float32
        r = rel(mine, ref)
        assert r <= 2e-5, (case, r)
        out[f"{case}/output"] = ref.numpy()
        print(f"{case:8s} {tuple(ref.shape)} oracle vs reference {r:.2e}, |out| {float(ref.norm()):.2f}")
    np.savez_compressed(GOLDEN / "upsampler.npz", **out)
    print(f"wrote {GOLDEN / 'upsampler.npz'}")
    return 0


if __name__ == "__main__":
    sys.exit(main())

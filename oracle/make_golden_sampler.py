"""ORACLE / test infrastructure only — generates tests/golden/sampler.npz by running the reference's OWN denoise loops
(mlx_video/generate.py: ``denoise_distilled`` :564-881, ``denoise_audio_only`` :888-1058, ``denoise_dev`` :1060-1327,
``denoise_dev_av`` :1330-1703; eager and, where the reference offers it, ``cfg_batch`` branches) unmodified over
oracle/mlx_shim, with the reference's own LTXModel and ``LatentState``, on the seeded cases of oracle/sampler_fixture.py.

    python oracle/make_golden_sampler.py     # needs /root/reference (this container); the fixture is committed

While generating it asserts that the oracle's loop restatements (ltx_oracle.denoise_*) reproduce every output
(<= 2e-5 relative: fp32 summation order inside the forward only) and that the reference's cfg_batch branch agrees
with its two-pass branch.
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))

import ltx_oracle as O  # noqa: E402,F401
import ref_loader  # noqa: E402
import sampler_fixture as SF  # noqa: E402
from make_golden import build_reference_model, ref_config, rel  # noqa: E402,F401

GOLDEN = HERE.parent / "tests" / "golden"


def run_reference(R, case: str, cfg_batch: bool = False):
    c, x = SF.CASES[case], SF.inputs(case)
    a = R.mx.array
    model = build_reference_model(R, SF.config(case), SF.weights(case))
    G = R.generate
    sig = SF.sigmas(case)
    state = None
    if c.get("state"):
        state = R.latent.LatentState(latent=a(x["latents"]), clean_latent=a(x["clean"]), denoise_mask=a(x["mask"]))
    if c["loop"] == "distilled":
        kw = {}
        if c.get("audio"):
            kw = dict(audio_latents=a(x["audio_latents"]), audio_positions=a(x["audio_positions"]), audio_embeddings=a(x["actx_pos"]))
        v, au = G.denoise_distilled(a(x["latents"]), a(x["positions"]), a(x["ctx_pos"]), model, [float(s) for s in sig],
                                    verbose=False, state=state, **kw)
        return v._t, None if au is None else au._t
    if c["loop"] == "dev":
        v = G.denoise_dev(a(x["latents"]), a(x["positions"]), a(x["ctx_pos"]), a(x["ctx_neg"]), model, a(torch.from_numpy(sig)),
                          cfg_scale=c["cfg_scale"], verbose=False, state=state, cfg_batch=cfg_batch)
        return v._t, None
    if c["loop"] == "dev_av":
        v, au = G.denoise_dev_av(a(x["latents"]), a(x["audio_latents"]), a(x["positions"]), a(x["audio_positions"]),
                                 a(x["ctx_pos"]), a(x["ctx_neg"]), a(x["actx_pos"]), a(x["actx_neg"]), model,
                                 a(torch.from_numpy(sig)), cfg_scale=c["cfg_scale"], verbose=False, video_state=state,
                                 cfg_batch=cfg_batch)
        return v._t, au._t
    if c["loop"] == "audio_only":
        au = G.denoise_audio_only(a(x["audio_latents"]), a(x["audio_positions"]), a(x["actx_pos"]), model,
                                  [float(s) for s in sig], verbose=False)
        return None, au._t
    raise KeyError(c["loop"])


def main() -> int:
    torch.set_num_threads(8)
    R = ref_loader.load()
    out = {}
    for case, c in SF.CASES.items():
        ref_v, ref_a = run_reference(R, case)
        mine_v, mine_a = SF.run_oracle(case)
        worst = 0.0
        for name, ref, mine in (("video", ref_v, mine_v), ("audio", ref_a, mine_a)):
            assert (ref is None) == (mine is None), (case, name)
            if ref is None:
                continue
            assert ref.dtype == torch.float32 and ref.shape == mine.shape
            r = rel(mine, ref)
            assert r <= 2e-5, (case, name, r)
            worst = max(worst, r)
            out[f"{case}/{name}"] = ref.numpy()
        note = ""
        if c["loop"] in ("dev", "dev_av") and c["cfg_scale"] != 1.0:  # the one-forward B=2 branch of the reference
            bv, ba = run_reference(R, case, cfg_batch=True)
            rb = max(rel(bv, ref_v), 0.0 if ba is None else rel(ba, ref_a))
            ov, oa = SF.run_oracle(case, cfg_batch=True)
            ro = max(rel(ov, bv), 0.0 if oa is None else rel(oa, ba))
            assert rb <= 2e-5 and ro <= 2e-5, (case, rb, ro)
            note = f", cfg_batch vs two-pass {rb:.1e}"
        moved = rel(ref_v, SF.inputs(case)["latents"]) if ref_v is not None else rel(ref_a, SF.inputs(case)["audio_latents"])
        print(f"{case:16s} oracle vs reference {worst:.2e}{note}; the loop moved the latents by {moved:.2f}")
    GOLDEN.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(GOLDEN / "sampler.npz", **out)
    print(f"wrote {GOLDEN / 'sampler.npz'}: {len(out)} arrays")
    return 0


if __name__ == "__main__":
    sys.exit(main())

"""GPU: the product's denoise loops (mlx_video_b200.sampler: denoise_distilled / denoise_audio_only / denoise_dev /
denoise_dev_av — forward through the C ABI, CFG + x0 + mask blend + fp32 Euler in ltxb_euler_step) against the latents
the reference's OWN loops produced (tests/golden/sampler.npz).  Tolerance as for the forward: rel-L2 <= 1e-2, cosine
>= 0.999 on the final latents."""
import pytest
import torch

import mlx_video_b200 as M
import sampler_fixture as SF
from mlx_video_b200 import sampler
from test_gpu_parity import assert_close, build

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def run_product(case: str, model, cfg_batch: bool = False):
    c, x = SF.CASES[case], {k: v.to(DEV) for k, v in SF.inputs(case).items()}
    sig = SF.sigmas(case)
    state = sampler.LatentState(x["latents"], x["clean"], x["mask"]) if c.get("state") else None
    if c["loop"] == "distilled":
        kw = {}
        if c.get("audio"):
            kw = dict(audio_latents=x["audio_latents"], audio_positions=x["audio_positions"], audio_embeddings=x["actx_pos"])
        return sampler.denoise_distilled(x["latents"], x["positions"], x["ctx_pos"], model, sig, state=state, **kw)
    if c["loop"] == "dev":
        return sampler.denoise_dev(x["latents"], x["positions"], x["ctx_pos"], x["ctx_neg"], model, sig,
                                   cfg_scale=c["cfg_scale"], state=state, cfg_batch=cfg_batch), None
    if c["loop"] == "dev_av":
        return sampler.denoise_dev_av(x["latents"], x["audio_latents"], x["positions"], x["audio_positions"], x["ctx_pos"],
                                      x["ctx_neg"], x["actx_pos"], x["actx_neg"], model, sig, cfg_scale=c["cfg_scale"],
                                      video_state=state, cfg_batch=cfg_batch)
    return None, sampler.denoise_audio_only(x["audio_latents"], x["audio_positions"], x["actx_pos"], model, sig)


@pytest.mark.parametrize("case", sorted(SF.CASES))
def test_loops_against_reference_golden(golden, case):
    g = golden("sampler")
    c = SF.CASES[case]
    model = build(SF.config(case), SF.weights(case))
    batched = (False, True) if c["loop"] in ("dev", "dev_av") and c["cfg_scale"] != 1.0 else (False,)
    for cfg_batch in batched:
        v, a = run_product(case, model, cfg_batch)
        torch.cuda.synchronize()
        model.check_timestep_groups()
        what = f"{case} cfg_batch={cfg_batch}"
        if f"{case}/video" in g:
            want = torch.from_numpy(g[f"{case}/video"])
            assert v.shape == want.shape and v.dtype == torch.float32
            assert_close(v, want, what + " video")
        else:
            assert v is None
        if f"{case}/audio" in g:
            want = torch.from_numpy(g[f"{case}/audio"])
            assert a.shape == want.shape and a.dtype == torch.float32
            assert_close(a, want, what + " audio")
        else:
            assert a is None


def test_loop_argument_errors():
    case = "distilled_av"
    model = build(SF.config(case), SF.weights(case))
    x = {k: v.to(DEV) for k, v in SF.inputs(case).items()}
    with pytest.raises(ValueError, match="audio_positions/audio_embeddings"):  # generate.py:628-629
        sampler.denoise_distilled(x["latents"], x["positions"], x["ctx_pos"], model, SF.sigmas(case), audio_latents=x["audio_latents"])

"""CPU: the oracle's restatement of the four denoise loops (ltx_oracle.denoise_*) against the latents the reference's
OWN loops produced over the shim (tests/golden/sampler.npz, oracle/make_golden_sampler.py).  The generator found them
bit-identical; the tolerance here only allows for another BLAS summation order on another host."""
import pytest
import torch

import sampler_fixture as SF
from conftest import rel_l2


@pytest.mark.parametrize("case", sorted(SF.CASES))
def test_oracle_loops_reproduce_reference(golden, case):
    g = golden("sampler")
    v, a = SF.run_oracle(case)
    assert (v is None) == (f"{case}/video" not in g) and (a is None) == (f"{case}/audio" not in g)
    if v is not None:
        assert rel_l2(v, torch.from_numpy(g[f"{case}/video"])) <= 2e-5
    if a is not None:
        assert rel_l2(a, torch.from_numpy(g[f"{case}/audio"])) <= 2e-5


def test_oracle_cfg_batch_branch_agrees():
    for case in ("dev", "dev_av_i2v"):
        v0, a0 = SF.run_oracle(case)
        v1, a1 = SF.run_oracle(case, cfg_batch=True)
        assert rel_l2(v1, v0) <= 2e-5 and (a0 is None or rel_l2(a1, a0) <= 2e-5)


def test_conditioned_frame_follows_the_mask(golden):
    """Hand check of the i2v cases: frame 0 has mask 0.25, so after the last step (latents = denoised, blended) it is
    0.25 * x0 + 0.75 * clean — it must sit closer to the clean latent than any free frame does."""
    g = golden("sampler")
    x = SF.inputs("distilled_i2v")
    out = torch.from_numpy(g["distilled_i2v/video"])
    d0 = (out[:, :, 0] - x["clean"][:, :, 0]).norm()
    d1 = (out[:, :, 1] - x["clean"][:, :, 1]).norm()
    assert d0 < 0.5 * d1

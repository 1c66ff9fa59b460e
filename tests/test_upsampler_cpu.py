"""CPU: the oracle's restatement of the latent upsampler (oracle/upsampler_oracle.py) against the outputs of the
reference's OWN upsampler.py run over the shim (tests/golden/upsampler.npz, oracle/make_golden_upsampler.py), plus hand
checks of the pieces.  No compute through the product here (it has no CPU path)."""
import pytest
import torch

import upsampler_fixture as UF
import upsampler_oracle as U
from conftest import rel_l2


@pytest.mark.parametrize("case", sorted(UF.CASES))
def test_oracle_reproduces_reference(golden, case):
    g = golden("upsampler")
    got = UF.run_oracle(case)
    want = torch.from_numpy(g[f"{case}/output"])
    assert got.shape == want.shape
    assert rel_l2(got, want) <= 2e-5  # fp32 summation order of the convolutions only


def test_pixel_shuffle_hand_check():
    """Channel (co, rh, rw) of pixel (h, w) lands at (2h + rh, 2w + rw, co) (upsampler.py:124-139)."""
    x = torch.arange(1 * 2 * 3 * 8, dtype=torch.float32).reshape(1, 2, 3, 8)  # Co = 2
    y = U.pixel_shuffle_cl(x)
    assert y.shape == (1, 4, 6, 2)
    for h, w, co, rh, rw in [(0, 0, 0, 0, 0), (1, 2, 1, 1, 0), (0, 1, 1, 0, 1), (1, 0, 0, 1, 1)]:
        assert y[0, 2 * h + rh, 2 * w + rw, co] == x[0, h, w, (co * 2 + rh) * 2 + rw]


def test_group_norm_statistics():
    """Per sample and group: zero mean, unit (population) variance before the affine; groups are independent."""
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 3, 4, 5, 64, generator=g) * 3 + 1
    y = U.group_norm_cl(x, torch.ones(64), torch.zeros(64))
    t = y.reshape(2, 60, 32, 2)
    assert float(t.mean(dim=(1, 3)).abs().max()) < 1e-5
    assert float((t.var(dim=(1, 3), correction=0) - 1).abs().max()) < 1e-3
    x2 = x.clone()
    x2[..., :2] += 100.0  # only group 0 changes
    y2 = U.group_norm_cl(x2, torch.ones(64), torch.zeros(64))
    assert torch.equal(y2[..., 2:], y[..., 2:])


def test_conv_layout_matches_sanitised_upstream_weights():
    """load_upsampler's layout mapping (upsampler.py:346-365) is the inverse of UF.upstream_state."""
    p = UF.params("b2_f1")
    back = U.sanitize_upsampler_weights(UF.upstream_state("b2_f1"))
    assert set(back) == set(p) and all(torch.equal(back[k], p[k]) for k in p)

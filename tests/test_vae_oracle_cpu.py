"""CPU: the oracle's restatement of the LTX-2 video VAE decoder (oracle/vae_decoder_oracle.py) against the outputs of the
reference's OWN LTX2VideoDecoder run over the shim (tests/golden/vae_decoder.npz, oracle/make_golden_vae.py), plus hand
checks of its pieces.  Groundwork for SURVEY.md §8f row N4, second half — there is no product path for it yet."""
import pytest
import torch

import vae_decoder_oracle as V
from conftest import rel_l2
from make_golden_vae import CASES, SEED, case_inputs


@pytest.fixture(scope="module")
def params():
    return V.init_decoder_params(SEED)


@pytest.mark.parametrize("case", sorted(CASES))
def test_oracle_reproduces_reference_decoder(golden, params, case):
    g = golden("vae_decoder")
    c = CASES[case]
    x = case_inputs(case)
    torch.manual_seed(1234)  # the generator's noise draw
    noise = torch.randn(x.shape)
    ts = None if c["timestep"] is None else torch.full((x.shape[0],), c["timestep"])
    got = V.decode(params, x, causal=c["causal"], timestep=ts, noise=noise, noise_scale=c["noise_scale"])
    sf, sh, sw = c["sample"]
    want = torch.from_numpy(g[f"{case}/output"])
    assert rel_l2(got[:, :, ::sf, ::sh, ::sw], want) <= 2e-5  # fp32 summation order of the convolutions only


def test_rearrangements_hand_check():
    """depth_to_space: channel (c, st, sh, sw) of voxel (d, h, w) lands at (2d+st, 2h+sh, 2w+sw) (sampling.py:143-157);
    unpatchify: channel (c, pr, pq) lands at (4h + pq, 4w + pr) — pr is the WIDTH offset (ops.py:47-80)."""
    x = torch.arange(1 * 16 * 2 * 2 * 3, dtype=torch.float32).reshape(1, 16, 2, 2, 3)  # c = 2
    y = V.depth_to_space(x)
    assert y.shape == (1, 2, 4, 4, 6)
    for c, st, sh, sw, d, h, w in [(0, 0, 0, 0, 0, 0, 0), (1, 1, 0, 1, 1, 1, 2), (0, 1, 1, 0, 0, 1, 1)]:
        assert y[0, c, 2 * d + st, 2 * h + sh, 2 * w + sw] == x[0, ((c * 2 + st) * 2 + sh) * 2 + sw, d, h, w]
    z = torch.arange(1 * 48 * 1 * 2 * 2, dtype=torch.float32).reshape(1, 48, 1, 2, 2)
    u = V.unpatchify(z)
    assert u.shape == (1, 3, 1, 8, 8)
    for c, pr, pq, h, w in [(0, 0, 0, 0, 0), (2, 3, 1, 1, 0), (1, 2, 3, 0, 1)]:
        assert u[0, c, 0, 4 * h + pq, 4 * w + pr] == z[0, (c * 4 + pr) * 4 + pq, 0, h, w]


def test_causal_convolution_sees_no_future_frames():
    g = torch.Generator().manual_seed(9)
    x = torch.randn(1, 8, 5, 4, 4, generator=g)
    w, b = torch.randn(6, 3, 3, 3, 8, generator=g) / 15, torch.randn(6, generator=g)
    y = V.causal_conv3d(x, w, b, causal=True)
    x2 = x.clone()
    x2[:, :, 3:] += 1.0  # change frames 3, 4
    y2 = V.causal_conv3d(x2, w, b, causal=True)
    assert y.shape == (1, 6, 5, 4, 4) and torch.equal(y2[:, :, :3], y[:, :, :3]) and not torch.equal(y2[:, :, 3:], y[:, :, 3:])
    assert not torch.equal(V.causal_conv3d(x2, w, b, causal=False)[:, :, 2], V.causal_conv3d(x, w, b, causal=False)[:, :, 2])


def test_pixel_norm_and_embedding():
    x = torch.randn(2, 16, 3, 4, 4)
    assert torch.allclose((V.pixel_norm(x) ** 2).mean(dim=1), torch.ones(2, 3, 4, 4), atol=1e-5)
    e = V.timestep_embedding_256(torch.tensor([0.0, 50.0]))
    assert e.shape == (2, 256) and torch.equal(e[0, :128], torch.ones(128)) and torch.equal(e[0, 128:], torch.zeros(128))  # [cos | sin]

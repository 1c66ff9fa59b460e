"""CPU: MLX affine-quantised checkpoints (SURVEY §8f rows N2 + N3).  The oracle's dequantised forward against the
velocities the reference's OWN ``LTXModel.from_pretrained`` + forward produced over the shim
(tests/golden/quant.npz, oracle/make_golden_quant.py), and the product's host-side ingest (key mapping, ``.scales``
detection, layout inference, quantization.json, load-time cast, ignored extras) on the same files.  No compute through
the product here (it has no CPU path)."""
import json

import numpy as np
import pytest
import torch

import ltx_oracle as O
import quant_fixture as QF
from conftest import rel_l2
from mlx_video_b200 import checkpoint as ck


@pytest.mark.parametrize("variant", sorted(QF.VARIANTS))
def test_oracle_reproduces_reference_on_quantised_checkpoint(golden, variant):
    g = golden("quant")
    state, dense = QF.build(variant)
    assert QF.packed_checksum(state) == int(g[f"{variant}/checksum"][0]), "the regenerated checkpoint drifted from the golden's"
    assert sum(k.endswith(".scales") for k in state) == int(g[f"{variant}/n_quantized"][0])
    got, _ = O.OracleLTXModel(QF.config(), dense)(QF.inputs(variant), None)
    assert rel_l2(got, torch.from_numpy(g[f"{variant}/velocity"])) <= 2e-5  # fp32 summation order only
    # runtime LoRA adapters on the quantised model (lora.py:188-217) == the same delta merged into the fp32 weights
    merged = O.apply_lora_to_weights(dense, [(QF.lora_state(variant), QF.LORA_STRENGTH)])
    got_l, _ = O.OracleLTXModel(QF.config(), merged)(QF.inputs(variant), None)
    assert rel_l2(got_l, torch.from_numpy(g[f"{variant}/velocity_lora"])) <= 2e-5
    assert rel_l2(got_l, got) > 2e-2


@pytest.mark.parametrize("bits", [2, 4, 8])
@pytest.mark.parametrize("group_size", [32, 64, 128])
def test_affine_pack_layout_and_round_trip(bits, group_size):
    """Hand checks of the published format: level j of a word sits in bits [bits*j, bits*(j+1)); scales * q + biases
    reproduces every group's larger-magnitude edge exactly and the rest to within one quantisation step."""
    g = torch.Generator().manual_seed(bits * 1000 + group_size)
    w = (torch.randn(24, 256, generator=g) / 16).to(torch.bfloat16)
    packed, s, b = O.affine_quantize(w, group_size, bits)
    assert packed.dtype == np.uint32 and packed.shape == (24, 256 * bits // 32) and s.shape == b.shape == (24, 256 // group_size)
    q = O.affine_unpack(packed, bits)
    per = 32 // bits
    word0 = int(packed[3, 1])
    assert [int(x) for x in q[3, per:2 * per]] == [(word0 >> (bits * j)) & ((1 << bits) - 1) for j in range(per)]
    assert q.min() >= 0 and q.max() <= (1 << bits) - 1
    d = O.affine_dequantize(packed, s, b, group_size, bits)
    grp = w.float().reshape(24, -1, group_size)
    edge = torch.where(grp.amin(-1).abs() > grp.amax(-1).abs(), grp.amin(-1), grp.amax(-1))
    dq = d.reshape(24, -1, group_size)
    hit = (dq - edge.unsqueeze(-1)).abs().amin(-1)
    assert float(hit.max()) <= 2.0 ** -8 * float(edge.abs().max())  # bias = edge (rounded to bf16)
    # within the refit's clipping (up to one step) + rounding (half a step) + the bf16 rounding of the scale (q * 2^-9 steps)
    assert float((d - w.float()).abs().max()) <= 2.5 * float(s.float().abs().max())
    assert O.quant_params_from_shapes(packed.shape[1], s.shape[1], 256) == (group_size, bits)


def test_ingest_reads_both_layouts(tmp_path):
    for variant, v in QF.VARIANTS.items():
        d = tmp_path / variant
        path = QF.write_checkpoint(variant, d)
        state, _ = QF.build(variant)
        got = ck.load_transformer_weights(path)
        extras = set(got) - set(state)
        assert extras == {"audio_patchify_proj.weight"}  # vae.* / embeddings connectors never belong to the transformer
        expected = set(state)
        got = ck.load_transformer_weights(path, expected=expected)
        assert set(got) == expected
        for k, want in state.items():
            if isinstance(want, np.ndarray):
                assert got[k].dtype == torch.uint32 and np.array_equal(got[k].numpy(), want), k
                s = got[k[: -len("weight")] + "scales"]
                in_f = s.shape[1] * v["group_size"]
                assert ck.quant_layout(k, tuple(got[k].shape), tuple(s.shape), (want.shape[0], in_f)) == (v["group_size"], v["bits"])
            else:
                assert got[k].dtype == torch.bfloat16 and torch.equal(got[k].float(), want.float()), k  # F32 on disk -> bf16 values
        meta = ck.read_quantization_meta(path)
        assert (meta.get("bits"), meta.get("group_size")) == ((8, 32) if v["meta"] else (None, None))


def test_quant_layout_errors():
    assert ck.quant_layout("w", (8, 64), (8, 8), (8, 512)) == (64, 4)
    assert ck.quant_layout("w", (8, 128), (8, 16), (8, 512), {"bits": 8, "group_size": 32, "mode": "affine"}) == (32, 8)
    with pytest.raises(ValueError, match="quantization.json"):
        ck.quant_layout("w", (8, 128), (8, 16), (8, 512), {"bits": 4, "group_size": 64})
    with pytest.raises(ValueError, match="mode"):
        ck.quant_layout("w", (8, 64), (8, 8), (8, 512), {"mode": "mxfp4"})
    with pytest.raises(ValueError, match="shape mismatch"):
        ck.quant_layout("w", (4, 64), (8, 8), (8, 512))
    with pytest.raises(ValueError, match="unsupported"):
        ck.quant_layout("w", (8, 48), (8, 8), (8, 512))  # 3 bits
    with pytest.raises(ValueError, match="unsupported"):
        ck.quant_layout("w", (8, 64), (8, 32), (8, 512))  # groups of 16


def test_sanitize_state_dict_and_cast():
    P = ck.PREFIX
    sd = {P + "transformer_blocks.0.attn1.to_out.0.weight": torch.randn(4, 4), P + "transformer_blocks.0.attn1.to_out.0.scales":
          torch.randn(4, 1), "vae.x": torch.zeros(1), P + "video_embeddings_connector.w": torch.zeros(1)}
    out = ck.sanitize_state_dict(sd)
    assert sorted(out) == ["transformer_blocks.0.attn1.to_out.scales", "transformer_blocks.0.attn1.to_out.weight"]
    assert out["transformer_blocks.0.attn1.to_out.weight"].dtype == torch.bfloat16      # ltx.py:613-615
    assert out["transformer_blocks.0.attn1.to_out.scales"].dtype == torch.float32       # ltx.py:606-609
    plain = ck.sanitize_state_dict({"patchify_proj.bias": torch.ones(3, dtype=torch.float16)})
    assert plain["patchify_proj.bias"].dtype == torch.float16  # only fp32 is cast


def test_broken_quantization_json_is_ignored(tmp_path):
    (tmp_path / "quantization.json").write_text("{not json")
    assert ck.read_quantization_meta(tmp_path / "m.safetensors") == {}
    (tmp_path / "quantization.json").write_text(json.dumps([1, 2]))
    assert ck.read_quantization_meta(tmp_path / "m.safetensors") == {}

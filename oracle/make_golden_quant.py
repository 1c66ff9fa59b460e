"""ORACLE / test infrastructure only — generates tests/golden/quant.npz by running the reference's OWN
``LTXModel.from_pretrained`` (mlx_video/models/ltx/ltx.py:535-885: header scan, key sanitising, ``.scales`` detection,
quantization.json, ``nn.quantize`` with the scales predicate, the streaming safetensors loader, the strict
missing-key check) and forward, unmodified, over oracle/mlx_shim on the two small MLX-quantised checkpoints of
oracle/quant_fixture.py.  What the shim restates is the MLX library part: ``nn.quantize`` / ``nn.QuantizedLinear`` /
``mx.quantized_matmul`` / ``mx.dequantize`` (published affine definition) and the Module parameter tree.

    python oracle/make_golden_quant.py      # needs /root/reference (this container); the fixture is committed

While generating it asserts that (i) the modules the reference quantised are exactly the ones the checkpoint names,
(ii) the parameters it ends up holding equal the oracle's dequantised state dict, (iii) ``OracleLTXModel`` on that
state dict reproduces the reference's velocity (<= 2e-5 relative: fp32 summation order only).
"""
from __future__ import annotations

import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))

import ltx_oracle as O  # noqa: E402
import quant_fixture as QF  # noqa: E402
import ref_loader  # noqa: E402
from make_golden import ref_config, ref_modality, rel  # noqa: E402

GOLDEN = HERE.parent / "tests" / "golden"


def main() -> int:
    torch.set_num_threads(8)
    R = ref_loader.load()
    import importlib

    from mlx.utils import tree_flatten

    lora_ref = importlib.import_module("mlx_video.lora")

    out = {}
    for variant, v in QF.VARIANTS.items():
        cfg = QF.config()
        state, dense = QF.build(variant)
        with tempfile.TemporaryDirectory() as td:
            path = QF.write_checkpoint(variant, Path(td))
            model = R.ltx.LTXModel.from_pretrained(path, ref_config(R, cfg), strict=True)
            # strict load of a file with a missing tensor must raise (ltx.py:874-881)
            broken = {k: val for k, val in state.items() if k != "transformer_blocks.1.ff.proj_out.bias"}
            entries = {(QF._upstream_name(k) if v["prefixed"] else k): QF._entry(val) for k, val in broken.items()}
            QF.write_safetensors(Path(td) / "broken.safetensors", entries)
            try:
                R.ltx.LTXModel.from_pretrained(Path(td) / "broken.safetensors", ref_config(R, cfg), strict=True)
                raise AssertionError("reference accepted a checkpoint with a missing parameter")
            except ValueError as e:
                assert "Missing 1 parameters" in str(e), e
        held = dict(tree_flatten(model.parameters()))
        q_mods = sorted(p for p, m in model.named_modules() if isinstance(m, R.nn.QuantizedLinear))
        want_mods = sorted(k[: -len(".scales")] for k in state if k.endswith(".scales"))
        assert q_mods == want_mods and q_mods, (len(q_mods), len(want_mods))
        assert all(m.group_size == v["group_size"] and m.bits == v["bits"] for _, m in model.named_modules()
                   if isinstance(m, R.nn.QuantizedLinear))
        assert set(held) == set(state), sorted(set(held) ^ set(state))[:8]
        for k, val in state.items():
            got = held[k]._t
            if isinstance(val, np.ndarray):
                assert np.array_equal(got.numpy().astype(np.uint32), val), k
            else:
                assert got.dtype == torch.bfloat16 and torch.equal(got.float(), val.float()), (k, got.dtype)
        m = QF.inputs(variant)
        ref_v, ref_a = model(video=ref_modality(R, m), audio=None)
        assert ref_a is None
        ref_v = ref_v._t.float()
        mine, _ = O.OracleLTXModel(cfg, dense)(m, None)
        r = rel(mine, ref_v)
        assert r <= 2e-5, (variant, r)
        out[f"{variant}/velocity"] = ref_v.numpy()
        # LoRA on the quantised model: the reference attaches runtime adapters (lora.py:219-275), y + (x A^T) B^T s
        with tempfile.TemporaryDirectory() as td:
            spec = lora_ref.LoraSpec(QF.write_lora(variant, Path(td)), QF.LORA_STRENGTH)
            lora_ref.apply_lora_to_model(model, [spec], verbose=True)
        adapters = sorted(p for p, mod in model.named_modules() if isinstance(mod, lora_ref.LoRAAdapter))
        assert adapters == sorted(name[: -len(".weight")] for _, name, _, _ in QF.LORA_TARGETS), adapters
        ref_l, _ = model(video=ref_modality(R, m), audio=None)
        ref_l = ref_l._t.float()
        merged = O.apply_lora_to_weights(dense, [(QF.lora_state(variant), QF.LORA_STRENGTH)])  # fp32: the same sum
        mine_l, _ = O.OracleLTXModel(cfg, merged)(m, None)
        r_l = rel(mine_l, ref_l)
        assert r_l <= 2e-5 and rel(ref_l, ref_v) > 1e-2, (variant, r_l, rel(ref_l, ref_v))
        out[f"{variant}/velocity_lora"] = ref_l.numpy()
        out[f"{variant}/checksum"] = np.asarray([QF.packed_checksum(state)], np.int64)
        out[f"{variant}/n_quantized"] = np.asarray([len(q_mods)], np.int64)
        print(f"{variant}: {len(q_mods)} quantised linears ({v['bits']} bits / group {v['group_size']}), "
              f"oracle vs reference rel {r:.2e} (with runtime LoRA {r_l:.2e}; LoRA moves the output by {rel(ref_l, ref_v):.2e}), "
              f"|v| {float(ref_v.norm()):.3f}")
    GOLDEN.mkdir(parents=True, exist_ok=True)
    np.savez_compressed(GOLDEN / "quant.npz", **out)
    print(f"wrote {GOLDEN / 'quant.npz'}")
    return 0


if __name__ == "__main__":
    sys.exit(main())

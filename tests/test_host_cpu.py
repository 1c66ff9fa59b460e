"""CPU: host-side logic of the product package and the C-ABI surface (no compute calls — there is no GPU
here).  The shared library must load without a driver and export every symbol include/ltxb.h declares."""
import ctypes
import re
from pathlib import Path

import numpy as np
import pytest
import torch

import mlx_video_b200 as M
from mlx_video_b200 import _lib, sampler

ROOT = Path(__file__).resolve().parent.parent


def header_symbols():
    text = (ROOT / "include" / "ltxb.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ltxb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    syms = header_symbols()
    assert len(syms) >= 17
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/ltxb.h but not exported by libltxb.so"
    assert sorted(_lib.SIGNATURES) == syms, "ctypes SIGNATURES and the header disagree"
    assert lib.ltxb_abi_version() >= 1


def test_no_cpu_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(M.LtxbError):
        M.LTXModel(M.production_config(num_layers=1))
    with pytest.raises(M.LtxbError):
        M.ops.cast_f32_to_bf16(torch.zeros(8), torch.zeros(8, dtype=torch.bfloat16))
    assert _lib.lib.ltxb_device_check() != 0 and "CUDA" in _lib.last_error()


def test_bad_arguments_are_rejected_before_launch():
    epi = _lib.Epilogue()
    rc = _lib.lib.ltxb_gemm_bf16(None, 64, None, 64, None, 64, 128, 128, 64, ctypes.byref(epi), 0, -1, None)
    assert rc == -1 and "null" in _lib.last_error()
    rc = _lib.lib.ltxb_attention_fwd(16, 64, 16, 64, 16, 64, 16, 64, 1, 8, 8, 1, 96, 1.0, None, None)
    assert rc == -2 and "head dim" in _lib.last_error()
    rc = _lib.lib.ltxb_euler_step(16, 16, None, 1.0, None, 0.0, 0.0, None, None, 4, 4, None, None)
    assert rc == -1 and "sigma" in _lib.last_error()


def test_config_mirrors_reference_fields():
    c = M.LTXModelConfig()
    assert c.positional_embedding_max_pos == [20, 2048, 2048] and c.audio_positional_embedding_max_pos == [20]
    assert c.inner_dim == 4096 and c.audio_inner_dim == 2048
    assert c.rope_type == M.LTXRopeType.INTERLEAVED and c.model_type == M.LTXModelType.AudioVideo  # dataclass defaults
    p = M.production_config()
    assert p.rope_type == M.LTXRopeType.SPLIT and p.double_precision_rope and p.num_layers == 48
    v = p.get_video_config()
    assert (v.dim, v.heads, v.d_head, v.context_dim) == (4096, 32, 128, 4096) and p.get_audio_config() is None
    d = M.LTXModelConfig.from_dict({"model_type": "ltx video only model", "rope_type": "split", "junk": 1})
    assert d.model_type == M.LTXModelType.VideoOnly and d.rope_type == M.LTXRopeType.SPLIT
    assert d.to_dict()["rope_type"] == "split"


def test_sampler_grids_and_schedules_bit_exact(golden):
    g = golden("grids")
    assert np.array_equal(sampler.create_position_grid(1, 5, 16, 16), g["g_1_5_16_16"])
    assert np.array_equal(sampler.create_position_grid(2, 3, 4, 6), g["g_2_3_4_6"])
    assert np.array_equal(sampler.create_position_grid(1, 9, 24, 24)[:, :, ::7, :], g["g_1_9_24_24"])
    assert np.array_equal(sampler.create_audio_position_grid(1, 68), g["a_1_68"])
    assert np.array_equal(sampler.create_audio_position_grid(2, 21), g["a_2_21"])
    for steps, ntok in [(40, 5184), (8, 1280), (30, None), (4, 320)]:
        assert np.array_equal(sampler.ltx2_scheduler(steps, ntok), g[f"sched_{steps}_{ntok}"])
    assert sampler.STAGE_1_SIGMAS == list(g["STAGE_1_SIGMAS"]) and sampler.STAGE_2_SIGMAS == list(g["STAGE_2_SIGMAS"])
    assert sampler.compute_audio_frames(65, 24.0) == 68


def test_rope_base_frequencies_match_oracle():
    import ltx_oracle as O
    from mlx_video_b200.rope import rope_base_frequencies
    for n_axes, dim in [(3, 4096), (1, 2048), (3, 512), (1, 256)]:
        assert torch.equal(rope_base_frequencies(10000.0, n_axes, dim), O.rope_freq_indices(10000.0, n_axes, dim))


def test_checkpoint_key_mapping_and_safetensors_roundtrip(tmp_path):
    """ltx.py:508-533 key renames; header scan; transformer-only filtering."""
    from safetensors.torch import save_file

    from mlx_video_b200 import checkpoint as ck

    P = "model.diffusion_model."
    assert ck.sanitize_key(P + "transformer_blocks.3.attn1.to_out.0.weight") == "transformer_blocks.3.attn1.to_out.weight"
    assert ck.sanitize_key(P + "transformer_blocks.0.ff.net.0.proj.bias") == "transformer_blocks.0.ff.proj_in.bias"
    assert ck.sanitize_key(P + "transformer_blocks.0.audio_ff.net.2.weight") == "transformer_blocks.0.audio_ff.proj_out.weight"
    assert ck.sanitize_key(P + "adaln_single.emb.timestep_embedder.linear_1.weight") == "adaln_single.emb.timestep_embedder.linear1.weight"
    assert ck.sanitize_key(P + "caption_projection.linear_2.bias") == "caption_projection.linear2.bias"
    assert ck.sanitize_key(P + "video_embeddings_connector.x.weight") is None and ck.sanitize_key("vae.decoder.conv.weight") is None
    assert ck.sanitize_key("patchify_proj.weight") == "patchify_proj.weight"  # already sanitised
    tensors = {P + "patchify_proj.weight": torch.randn(8, 4).bfloat16(), P + "transformer_blocks.0.attn1.to_out.0.bias": torch.randn(8),
               "vae.decoder.w": torch.zeros(2), P + "audio_embeddings_connector.w": torch.zeros(2)}
    save_file(tensors, str(tmp_path / "m.safetensors"))
    assert sorted(ck.scan_keys([tmp_path / "m.safetensors"])) == sorted(tensors)
    got = ck.load_transformer_weights(tmp_path)
    assert sorted(got) == ["patchify_proj.weight", "transformer_blocks.0.attn1.to_out.bias"]
    assert torch.equal(got["patchify_proj.weight"], tensors[P + "patchify_proj.weight"])
    # in an upstream-layout file only prefixed tensors are the transformer's (ltx.py:548-553)
    save_file({P + "patchify_proj.bias": torch.zeros(2), "patchify_proj.weight": torch.zeros(2, 2)}, str(tmp_path / "u.safetensors"))
    assert sorted(ck.load_transformer_weights(tmp_path / "u.safetensors")) == ["patchify_proj.bias"]
    with pytest.raises(FileNotFoundError):
        ck.load_transformer_weights(tmp_path / "empty_dir_that_does_not_exist")


def test_model_sanitize_matches_reference_key_mapping():
    """LTXModel.sanitize (ltx.py:508-533) needs no device: unbound call on a bare object."""
    from mlx_video_b200.model import LTXModel

    P = "model.diffusion_model."
    w = {P + "transformer_blocks.0.attn1.to_out.0.weight": 1, P + "transformer_blocks.0.ff.net.2.bias": 2,
         P + "caption_projection.linear_1.weight": 3, P + "audio_embeddings_connector.x": 4, "vae.decoder.w": 5,
         "patchify_proj.weight": 6}
    got = LTXModel.sanitize(None, w)
    assert got == {"transformer_blocks.0.attn1.to_out.weight": 1, "transformer_blocks.0.ff.proj_out.bias": 2,
                   "caption_projection.linear1.weight": 3}

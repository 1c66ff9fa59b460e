"""ctypes binding of ``libltxb.so`` (C ABI declared in ``include/ltxb.h``).

There is no fallback: if the shared library is missing this module raises at import time, and every
wrapper raises ``LtxbError`` when a call returns non-zero.  The product path never routes through
``oracle/`` or a PyTorch eager implementation.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

_HERE = Path(__file__).resolve().parent
CSRC = _HERE / "csrc"
LIB_PATH = Path(os.environ["LTXB_LIB"]) if os.environ.get("LTXB_LIB") else CSRC / "libltxb.so"  # LTXB_LIB: A/B a build

# (name, restype, argtypes) — must list EVERY symbol declared in include/ltxb.h (tests check this).
_i32, _i64, _f32, _vp = C.c_int32, C.c_int64, C.c_float, C.c_void_p


ABI_VERSION = 4  # include/ltxb.h: ltxb_abi_version()


class PeerSync(C.Structure):
    """Mirror of ``struct ltxb_peer_sync`` (the folded cross-GPU flag barrier)."""

    _fields_ = [
        ("flags", _vp * 8),
        ("n_peers", _i32),
        ("my_rank", _i32),
        ("epoch_counter", _vp),
        ("done_counter", _vp),
    ]


class Epilogue(C.Structure):
    """Mirror of ``struct ltxb_epilogue``."""

    _fields_ = [
        ("mode", _i32),
        ("gate_row_div", _i32),
        ("bias", _vp),
        ("resid", _vp),
        ("ldr", _i64),
        ("gate", _vp),
        ("gate_ld", _i64),
        ("gate_row_index", _vp),
        ("gate_table", _vp),
        ("a_group_cols", _i32),
        ("flags", _i32),
        ("a_group_stride", _i64),
        ("peer_sync", C.POINTER(PeerSync)),
    ]


EPI_BIAS_BF16, EPI_GELU_BF16, EPI_SILU_BF16, EPI_BIAS_F32, EPI_RESID_GATE_F32 = range(5)

SIGNATURES = {
    "ltxb_last_error": (C.c_char_p, []),
    "ltxb_abi_version": (C.c_int, []),
    "ltxb_device_check": (C.c_int, []),
    "ltxb_set_device": (C.c_int, [_i32]),
    "ltxb_kernel_launches": (C.c_int64, []),
    "ltxb_gemm_bf16": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _i32, C.POINTER(Epilogue), _i32, _i32, _vp]),
    "ltxb_gemm_qw_bf16": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _i64, _i32, _i32, _i32, _vp, _i32, _vp]),
    "ltxb_gemm_workspace_bytes": (C.c_int64, []),
    "ltxb_gemm_set_workspace": (C.c_int, [_vp, _i64, _vp]),
    "ltxb_rmsnorm_modulate": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _i32, _f32, _vp, _i64, _i32, _i32, _vp, _vp, _i32, _vp, _vp]),
    "ltxb_residual_rmsnorm_modulate": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _f32, _vp, _i64, _i32, _i32, _i32, _vp, _vp, _vp, _i32, _vp, _vp]),
    "ltxb_layernorm_modulate": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _i32, _f32, _vp, _i64, _vp, _vp, _i32, _vp, _vp]),
    "ltxb_gate_residual": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _i32, _vp, _i64, _i32, _vp, _i32, _vp, _vp]),
    "ltxb_qknorm_rope": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _i32, _vp, _f32, _vp, _vp, _i32, _vp]),
    "ltxb_qknorm_rope_segments": (C.c_int, [_vp, _i64, _i32, _i64, _i32, _i32, _i32, _i32, _vp, _i64, _vp, _f32, _vp, _vp, _i32, _vp]),
    "ltxb_qknorm_rope_scatter": (C.c_int, [_vp, _i64, _vp, _i64, _i32, _i64, _i32, _i32, _i32, _i32, _vp, _f32, _vp, _vp, _i32, _vp]),
    "ltxb_qknorm_rope_scatter_peers": (C.c_int, [_vp, _i64, C.POINTER(_vp), _i32, _i64, _i32, _i32, _i32, _i32, _vp, _f32, _vp, _vp, _i32, _vp]),
    "ltxb_qkv_norm_rope_scatter_peers": (C.c_int, [_vp, _i64, _i64, C.POINTER(_vp), _i32, _i64, _i64, _i32, _i32, _i32, _vp, _f32, _vp, _vp, _i32, _vp]),
    "ltxb_peer_barrier": (C.c_int, [C.POINTER(_vp), _i32, _i32, _vp, _vp]),
    "ltxb_peer_broadcast": (C.c_int, [_vp, _i64, C.POINTER(_vp), _i32, _vp]),
    "ltxb_attention_fwd_peers_sync": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, C.POINTER(_vp), _i32, _i32, _i64, _i32, _i32, _i32, _i32, _f32, C.POINTER(PeerSync), _vp]),
    "ltxb_attention_fwd_peers": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, C.POINTER(_vp), _i32, _i32, _i64, _i32, _i32, _i32, _i32, _f32, _vp]),
    "ltxb_attention_partial_floats": (C.c_int64, [_i32, _i32, _i32, _i32]),
    "ltxb_attention_partial": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i32, _i32, _i32, _i32, _i32, _f32, _vp]),
    "ltxb_attention_merge": (C.c_int, [_vp, _i64, _i32, _vp, _i64, _i32, _i32, _i32, _i32, _vp]),
    "ltxb_timestep_embed": (C.c_int, [_vp, _i32, _f32, _i32, _vp, _i64, _vp]),
    "ltxb_timestep_groups": (C.c_int, [_vp, _i32, _i32, _vp, _vp, _vp, _vp]),
    "ltxb_rope_table": (C.c_int, [_vp, _i32, _i32, _i32, C.POINTER(_f32), _vp, _i32, _i32, _i32, _i32, _vp, _vp, _vp]),
    "ltxb_silu_bf16": (C.c_int, [_vp, _vp, _i64, _vp]),
    "ltxb_cast_f32_to_bf16": (C.c_int, [_vp, _vp, _i64, _vp]),
    "ltxb_cast_bf16_to_f32": (C.c_int, [_vp, _vp, _i64, _vp]),
    "ltxb_attention_fwd": (C.c_int, [_vp, _i64, _vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _f32, _vp, _vp]),
    "ltxb_attention_workspace_bytes": (C.c_int64, []),
    "ltxb_attention_set_workspace": (C.c_int, [_vp, _i64]),
    "ltxb_lora_merge_bf16": (C.c_int, [_vp, _i64, _vp, _i64, _i64, _i32, _f32, _vp]),
    "ltxb_dequant_affine_bf16": (C.c_int, [_vp, _i64, _vp, _vp, _i64, _i32, _vp, _i64, _i64, _i32, _i32, _i32, _vp]),
    "ltxb_im2col_cl": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "ltxb_groupnorm_workspace_bytes": (C.c_int64, [_i32, _i64, _i32]),
    "ltxb_groupnorm_silu": (C.c_int, [_vp, _vp, _i32, _i64, _i32, _i32, _f32, _vp, _vp, _vp, _i32, _vp, _i64, _vp]),
    "ltxb_pixel_shuffle2": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _vp]),
    "ltxb_latent_layout": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i64, _i32, _vp]),
    "ltxb_vae_gather_rows": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i64, _i64, _vp, _vp, _vp, _vp, _i64, _f32, _i32, _vp]),
    "ltxb_vae_depth_to_space": (C.c_int, [_vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _vp]),
    "ltxb_vae_prepare_latent": (C.c_int, [_vp, _vp, _f32, _vp, _vp, _vp, _i64, _i32, _i64, _vp]),
    "ltxb_vae_unpatchify": (C.c_int, [_vp, _vp, _i64, _i32, _i32, _i32, _vp]),
    "ltxb_vae_blend_tile": (C.c_int, [_vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "ltxb_vae_blend_normalize": (C.c_int, [_vp, _vp, _i64, _i64, _vp]),
    "ltxb_euler_step": (C.c_int, [_vp, _vp, _vp, _f32, _vp, _f32, _f32, _vp, _vp, _i64, _i32, _vp, _vp]),
}


class LtxbError(RuntimeError):
    """A C-ABI call returned a non-zero status."""


def build(verbose: bool = False) -> Path:
    """Compile ``libltxb.so`` in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
    cmd = ["make", "-C", str(CSRC), "-j4"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"building libltxb.so failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stdout)
    return LIB_PATH


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        if os.environ.get("LTXB_AUTOBUILD", "1") == "1":
            build()
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} is missing: run `make -C {CSRC}` (or __graft_entry__.build()). "
                "There is no CPU / eager fallback for the LTX-2 DiT path."
            )
    lib = C.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch: fail loudly
        fn.restype = restype
        fn.argtypes = argtypes
    if lib.ltxb_abi_version() != ABI_VERSION:  # struct layouts below (Epilogue, PeerSync) belong to exactly this version
        raise ImportError(f"{LIB_PATH} has ABI version {lib.ltxb_abi_version()}, this binding is for {ABI_VERSION}: rebuild it (make -C {CSRC})")
    return lib


lib = _load()


def last_error() -> str:
    return (lib.ltxb_last_error() or b"").decode()


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise LtxbError(f"{what} failed with status {rc}: {last_error()}")

"""ORACLE / test infrastructure only.

Imports the reference's own hot-path modules from /root/reference UNMODIFIED, with the third-party
``mlx`` dependency replaced by ``oracle/mlx_shim`` and the reference's package ``__init__`` files
(which pull in the whole product: converters, VAEs, text encoder) bypassed by registering bare
package stubs.  Only usable where /root/reference exists (this container — not the GPU box).
"""
from __future__ import annotations

import importlib
import sys
import types
from pathlib import Path

REFERENCE_ROOT = Path("/root/reference")
_SHIM = Path(__file__).resolve().parent / "mlx_shim"


def available() -> bool:
    return (REFERENCE_ROOT / "mlx_video" / "models" / "ltx" / "ltx.py").exists()


def _stub_package(name: str, path: Path) -> None:
    if name in sys.modules:
        return
    mod = types.ModuleType(name)
    mod.__path__ = [str(path)]  # namespace-style: submodules resolve, __init__.py is NOT executed
    mod.__package__ = name
    sys.modules[name] = mod


def load():
    """Returns a namespace with the reference modules: ltx, transformer, attention, rope, adaln,
    feed_forward, text_projection, config, utils, generate."""
    if not available():
        raise RuntimeError("reference sources not present at /root/reference")
    if str(_SHIM) not in sys.path:
        sys.path.insert(0, str(_SHIM))
    import mlx.core  # noqa: F401  (the shim)

    assert "mlx_shim" in mlx.core.__file__, "a real mlx is installed; the shim must not shadow it silently"
    pkg = REFERENCE_ROOT / "mlx_video"
    _stub_package("mlx_video", pkg)
    _stub_package("mlx_video.models", pkg / "models")
    _stub_package("mlx_video.models.ltx", pkg / "models" / "ltx")
    ns = types.SimpleNamespace()
    ns.mx = importlib.import_module("mlx.core")
    ns.nn = importlib.import_module("mlx.nn")
    for short, full in [
        ("config", "mlx_video.models.ltx.config"),
        ("utils", "mlx_video.utils"),
        ("rope", "mlx_video.models.ltx.rope"),
        ("adaln", "mlx_video.models.ltx.adaln"),
        ("attention", "mlx_video.models.ltx.attention"),
        ("feed_forward", "mlx_video.models.ltx.feed_forward"),
        ("text_projection", "mlx_video.models.ltx.text_projection"),
        ("transformer", "mlx_video.models.ltx.transformer"),
        ("ltx", "mlx_video.models.ltx.ltx"),
        ("latent", "mlx_video.conditioning.latent"),
        ("generate", "mlx_video.generate"),
    ]:
        setattr(ns, short, importlib.import_module(full))
    return ns


def set_param(model, name: str, value) -> None:
    """Assign a flat state-dict entry (reference naming, ltx.py:508-533) onto the reference model."""
    obj = model
    parts = name.split(".")
    for p in parts[:-1]:
        if isinstance(obj, dict):
            obj = obj[int(p)] if p.isdigit() else obj[p]
        else:
            obj = getattr(obj, p)
    if isinstance(obj, dict):
        obj[parts[-1]] = value
    else:
        if not hasattr(obj, parts[-1]):
            raise KeyError(f"reference model has no parameter {name}")
        setattr(obj, parts[-1], value)

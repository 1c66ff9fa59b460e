"""ORACLE / test infrastructure only — a minimal stand-in for the third-party ``mlx`` package
(mlx==0.30.1 in the reference's uv.lock; Apple-only, not installable here), backed by torch on CPU.

It exists so that the reference's OWN source files (/root/reference/mlx_video/models/ltx/*.py and the
sampler helpers of mlx_video/generate.py) can be imported and executed unmodified to produce golden
vectors (oracle/make_golden.py).  Only the API surface those files touch is provided; every op follows
MLX's published semantics.  Never imported by the product package.
"""

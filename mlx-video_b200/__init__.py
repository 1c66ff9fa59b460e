"""B200-native LTX-2 DiT denoise-step forward (drop-in for mlx_video.models.ltx.LTXModel).

The directory is named ``mlx-video_b200`` (not importable as such); import it as ``mlx_video_b200``
through the loader module of that name at the repository root.
"""
__version__ = "0.1.0"

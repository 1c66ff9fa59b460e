"""CPU: host logic of the packed-weight registry (mlx-video_b200/packed.py) and the C struct layouts the ctypes binding mirrors."""
import ctypes as C
import subprocess
import sys
from pathlib import Path

import torch

import mlx_video_b200  # noqa: F401
from mlx_video_b200 import _lib
from mlx_video_b200.packed import PackedWeights

ROOT = Path(__file__).resolve().parent.parent


def _fake(rows, cols, bits, group):
    return (torch.zeros(rows, cols * bits // 32, dtype=torch.int32), torch.zeros(rows, cols // group, dtype=torch.bfloat16),
            torch.zeros(rows, cols // group, dtype=torch.bfloat16))


def test_registry_merges_adjacent_slices_and_resolves_row_ranges():
    reg = PackedWeights()
    fused = torch.zeros(3 * 64, 256, dtype=torch.bfloat16)  # q | k | v rows of one fused matrix
    other = torch.zeros(64, 256, dtype=torch.bfloat16)
    for i in range(3):
        p, s, b = _fake(64, 256, 4, 64)
        p += i + 1  # tell the slices apart
        reg.register(fused[64 * i:64 * (i + 1)], p, s, b, 64, 4)
    p, s, b = _fake(64, 256, 8, 32)
    reg.register(other, p, s, b, 32, 8)
    assert len(reg) == 4
    reg.merge_adjacent()
    assert len(reg) == 2
    hit = reg.lookup(fused)  # the fused view
    assert hit is not None and hit[0].shape == (192, 32) and hit[3:] == (64, 4)
    assert hit[0][0, 0] == 1 and hit[0][64, 0] == 2 and hit[0][128, 0] == 3
    kv = reg.lookup(fused[64:])  # k | v rows (what a cross-attention context projection uses)
    assert kv is not None and kv[0].shape == (128, 32) and kv[0][0, 0] == 2
    assert reg.lookup(fused[:, :128]) is None  # a column slice is not a row range of the registered matrix
    assert reg.lookup(torch.zeros(64, 256, dtype=torch.bfloat16)) is None
    o = reg.lookup(other)
    assert o is not None and o[3:] == (32, 8) and o[0].shape == (64, 64)
    p2, s2, b2 = _fake(64, 256, 2, 64)
    reg.register(torch.zeros(64, 256, dtype=torch.bfloat16), p2, s2, b2, 64, 2)  # 2-bit weights stay on the expanded copy
    assert len(reg) == 2
    reg.clear()
    assert len(reg) == 0 and reg.lookup(fused) is None


def test_ctypes_struct_layouts_match_the_header(tmp_path):
    """sizeof / offsetof of ltxb_epilogue and ltxb_peer_sync as gcc sees include/ltxb.h == what _lib.Epilogue / _lib.PeerSync declare."""
    src = tmp_path / "layout.c"
    fields_e = [f[0] for f in _lib.Epilogue._fields_]
    fields_p = [f[0] for f in _lib.PeerSync._fields_]
    body = "".join(f'printf("e.{f} %zu\\n", offsetof(ltxb_epilogue, {f}));' for f in fields_e)
    body += "".join(f'printf("p.{f} %zu\\n", offsetof(ltxb_peer_sync, {f}));' for f in fields_p)
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "ltxb.h"\nint main(void) {'
                   'printf("e %zu\\np %zu\\n", sizeof(ltxb_epilogue), sizeof(ltxb_peer_sync));' + body + "return 0; }\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", str(ROOT / "include"), str(src), "-o", str(exe)], check=True)
    got = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    assert int(got["e"]) == C.sizeof(_lib.Epilogue) and int(got["p"]) == C.sizeof(_lib.PeerSync)
    for f in fields_e:
        assert int(got[f"e.{f}"]) == getattr(_lib.Epilogue, f).offset, f
    for f in fields_p:
        assert int(got[f"p.{f}"]) == getattr(_lib.PeerSync, f).offset, f
    assert sys.maxsize > 2 ** 32

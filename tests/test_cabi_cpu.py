"""CPU: the C-ABI library builds for sm_100a, loads without a GPU / libcuda, and exports every
symbol include/dkt_stereo_b200.h declares.  No compute calls here."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "dkt_stereo_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dkt_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    from dkt_stereo_b200 import _lib
    lib = _lib.load()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
    # the ctypes table covers the header one to one
    assert sorted(_lib.SIGNATURES) == names
    assert lib.dkt_abi_version() == _lib.ABI_VERSION == 3
    assert lib.dkt_split_format() in (0, 1)
    assert lib.dkt_error_string(-2).decode().startswith("shape or option")


def test_struct_layout_matches_header(tmp_path):
    """sizeof / offsetof of the two ABI structs as gcc lays them out from the header == the ctypes mirror."""
    import subprocess
    from dkt_stereo_b200 import _lib
    fields_t = [n for n, _ in _lib.DktTensor._fields_]
    fields_e = [n for n, _ in _lib.DktEpilogue._fields_]
    src = tmp_path / "layout.c"
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT}/include/dkt_stereo_b200.h"', 'int main(void){',
             'printf("%zu %zu\\n", sizeof(dkt_tensor), sizeof(dkt_epilogue));']
    lines += [f'printf("%zu\\n", offsetof(dkt_tensor, {f}));' for f in fields_t]
    lines += [f'printf("%zu\\n", offsetof(dkt_epilogue, {f}));' for f in fields_e]
    lines += ['return 0;}']
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-std=c99", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    vals = [int(v) for v in out]
    assert vals[0] == ctypes.sizeof(_lib.DktTensor) and vals[1] == ctypes.sizeof(_lib.DktEpilogue)
    mine = [getattr(_lib.DktTensor, f).offset for f in fields_t] + [getattr(_lib.DktEpilogue, f).offset for f in fields_e]
    assert vals[2:] == mine


def test_argument_validation_without_gpu():
    """Invalid arguments are rejected before any CUDA call (so this is safe on a CPU box)."""
    from dkt_stereo_b200 import _lib
    lib = _lib.load()
    assert lib.dkt_corr1d_build_f32(None, None, 0, 0, 0, 0, None, 1, 1, 1, 1, 1, 1, 1.0, None) == -1
    assert lib.dkt_convex_upsample(None, 2, None, None, 1, 1, 1, 4, None) == -1
    assert lib.dkt_conv2d_tc(None, 1, None, None, 3, 64, None, 1, 8, 8, None) == -1
    # IGEV pre-loop volume kernels: null pointers / bad sizes -> DKT_E_INVALID, unsupported shapes -> DKT_E_UNSUPPORTED
    assert lib.dkt_gwc_volume(None, None, None, 1, 96, 8, 48, 8, 8, None) == -1
    assert lib.dkt_gwc_volume(1 << 12, 1 << 13, 1 << 14, 1, 96, 7, 48, 8, 8, None) == -1        # 96 % 7 != 0
    assert lib.dkt_gwc_volume(1 << 12, 1 << 13, 1 << 14, 1, 256, 8, 48, 8, 8, None) == -2      # 32 channels per group
    assert lib.dkt_conv3d_k3(None, None, None, None, None, 1.0, None, 1, 8, 8, 8, 8, 8, 1, None) == -1
    assert lib.dkt_conv3d_k3(1 << 12, 1 << 13, None, None, None, 1.0, 1 << 14, 1, 8, 8, 8, 8, 8, 3, None) == -2   # stride 3
    assert lib.dkt_conv3d_k3(1 << 12, 1 << 13, None, None, None, 1.0, 1 << 12, 1, 8, 8, 8, 8, 8, 1, None) == -1   # in == out
    assert lib.dkt_conv3d_c8(1 << 12, 1 << 13, None, None, None, 1.0, 1 << 14, 1, 4, 8, 8, 8, None) == -2         # CO not 8 | 1
    assert lib.dkt_deconv3d_k4s2(None, None, None, None, 1.0, None, 1, 16, 8, 4, 4, 4, None) == -1
    assert lib.dkt_deconv3d_k4s2(1 << 12, 1 << 13, None, None, 1.0, 1 << 14, 1, 6, 8, 4, 4, 4, None) == -2        # CI % 4
    assert lib.dkt_conv3d_k1(None, 16, None, 0, None, None, None, None, 1.0, None, 1, 16, 4, 4, 4, None) == -1
    assert lib.dkt_conv3d_k1(1 << 12, 100, 1 << 13, 100, 1 << 14, None, None, None, 1.0, 1 << 15, 1, 16, 4, 4, 4, None) == -2
    assert lib.dkt_softargmin(None, None, 1, 48, 8, 8, None) == -1


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "dkt_stereo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle/make_golden.py", ""), f"{f} mentions the oracle"
    for f in ("tools/evaluate_stereo.py",):
        path = os.path.join(ROOT, f)
        if os.path.exists(path):
            assert "import oracle" not in open(path).read() and "from oracle" not in open(path).read()

"""CPU suite, part 3: the C-ABI library builds for sm_100a, loads, exports every symbol that
include/*.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import glob
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    names = []
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = open(h).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"\b(xq_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def test_build_and_exports():
    from cn_chess_ai_b200 import build
    lib = build.build_native()
    assert os.path.exists(lib)
    out = subprocess.run(["nm", "-D", "--defined-only", lib], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (xq_[a-z0-9_]+)", out))
    want = declared_functions()
    assert len(want) >= 20
    missing = [n for n in want if n not in exported]
    assert not missing, f"declared in include/*.h but not exported: {missing}"
    L = C.CDLL(lib)
    for n in want:
        assert getattr(L, n) is not None


def test_sass_is_sm100a():
    from cn_chess_ai_b200 import build
    lib = build.build_native()
    out = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import cn_chess_ai_b200 as xq
    with pytest.raises(xq.XQError):
        xq.BatchedEnv(4)


def test_product_never_imports_oracle():
    for path in glob.glob(os.path.join(ROOT, "cn_chess_ai_b200", "**", "*"), recursive=True):
        if os.path.isfile(path) and path.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
            text = open(path).read()
            assert "xq_oracle" not in text and "from oracle" not in text and "import oracle" not in text, path
            assert "libxq_ref" not in text, path

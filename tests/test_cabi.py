"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
that include/*.h declares (no compute calls: there is no GPU here), and the product refuses to
run without a device instead of falling back to a CPU path."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "april_ann_b200", "libb200ann.so")


def declared_symbols():
    names = []
    for hdr in ("b200ann.h", "b200ann_host.h"):
        text = open(os.path.join(ROOT, "include", hdr)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names += re.findall(r"\b(b200h?_[a-z0-9_]+)\s*\(", text)
    return sorted(set(names))


def ensure_built():
    if not os.path.exists(LIB):
        import __graft_entry__ as g
        g.build()
    return LIB


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(ensure_built())
    syms = declared_symbols()
    assert len(syms) > 80
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_no_cpu_fallback_without_device():
    ensure_built()
    import april_ann_b200 as ann
    if ann.is_cuda_available():
        pytest.skip("a device is present")
    with pytest.raises(ann.B200Error):
        ann.Context(0)
    # the host-side objects that need no device still work and mirror the reference
    r = ann.random(1234)
    from oracle import MTRand
    o = MTRand(1234)
    assert [r.randInt() for _ in range(5)] == [o.randInt32() for _ in range(5)]
    assert list(ann.random(5678).shuffle(800)) == MTRand(5678).shuffle(800)
    assert abs(ann.random(7).rand(2.0) - MTRand(7).rand(2.0)) == 0.0


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under april_ann_b200/ may reference it."""
    pkg = os.path.join(ROOT, "april_ann_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "oracle/" not in src and "oracle import" not in src, f
    # tools/ are measurement helpers of the product: they do not use the checker either (the two scripts that
    # compare against it live under tests/)
    for f in os.listdir(os.path.join(ROOT, "tools")):
        if f.endswith(".py"):
            src = open(os.path.join(ROOT, "tools", f), errors="replace").read()
            assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f

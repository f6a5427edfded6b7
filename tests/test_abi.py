"""CPU-side checks of the drop-in boundary: libgsi.so loads and exports every symbol that
include/gsi.h declares; without a device the entry points fail loudly (no CPU fallback)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "gsi.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gsi_[a-z0-9_]+)\s*\(", text)) - {"gsi_record_sink"})


def test_library_exports_every_declared_symbol():
    from collaborative_filtering_b200 import _lib
    if not os.path.exists(_lib.SO_PATH):
        _lib.build()
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), "libgsi.so does not export %s" % name
    assert sorted(_lib.SYMBOLS) == declared, "python binding table out of sync with include/gsi.h"
    assert b"sm_100a" in lib.gsi_version()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from collaborative_filtering_b200.api import Context, GsiError
    with pytest.raises(GsiError) as e:
        Context(0)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def test_product_never_touches_the_checker():
    pkg = os.path.join(ROOT, "collaborative_filtering_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text, "%s references oracle/" % f

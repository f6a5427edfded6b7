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


def test_header_is_plain_c99(tmp_path):
    """The boundary is a C ABI: include/gsi.h must compile as C99 on its own (no C++ or CUDA types in the signatures),
    and a C translation unit that calls an entry point must link against libgsi.so."""
    import subprocess
    from collaborative_filtering_b200 import _lib
    hdr = os.path.join(ROOT, "include", "gsi.h")
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-xc", hdr])
    src = tmp_path / "probe.c"
    src.write_text('#include <stdio.h>\n#include "gsi.h"\n'
                   'int main(void) { gsi_ctx* c = 0; int rc = gsi_create(&c, 0, 0);\n'
                   '  printf("%s|%d|%s\\n", gsi_version(), rc, gsi_last_error(c));\n'
                   '  if (c) gsi_destroy(c);\n  return 0; }\n')
    exe = str(tmp_path / "probe")
    if not os.path.exists(_lib.SO_PATH):
        _lib.build()
    libdir = os.path.dirname(_lib.SO_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", exe, str(src), "-L", libdir, "-lgsi",
                           "-Wl,-rpath," + libdir])
    out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "sm_100a" in out.stdout
    import torch
    if not torch.cuda.is_available():
        assert "|2|" in out.stdout and "no CPU fallback" in out.stdout      # GSI_ERR_CUDA, loudly

"""CPU check of the CLI hosts' text primitives (collaborative_filtering_b200/csrc/host/fast_fmt.hpp): the fast
`%g` writer and the fast decimal reader must be byte / bit identical to libc's printf("%g") and strtod, which is
what the reference's `ostream << double` and `istream >> double` do (precompute_local.cpp:263-280,
local_calc_precomp.cpp:425-477).  The C++ harness compares millions of values of the classes the tools print."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fast_format_and_parse_match_libc(tmp_path):
    exe = str(tmp_path / "test_fast_fmt")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "host", "test_fast_fmt.cpp")])
    out = subprocess.run([exe, "200000"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "mismatches 0" in out.stdout

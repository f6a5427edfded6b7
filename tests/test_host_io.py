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


def test_local_calc_cli_argument_errors_and_no_cpu_fallback(tmp_path):
    """local_calc.cpp:570-581: a command line that does not parse prints the reference's message and exits with
    EXIT_FAILURE; without a CUDA device the tool stops at gsi_create -- there is no CPU path behind it."""
    import torch
    exe = os.path.join(ROOT, "collaborative_filtering_b200", "bin", "local_calc")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "collaborative_filtering_b200", "csrc")], stdout=subprocess.DEVNULL)
    p = subprocess.run([exe, "--bogus"], cwd=str(tmp_path), capture_output=True, text=True, timeout=60)
    assert p.returncode == 1 and "Error in parsing command line arguments." in p.stdout
    if torch.cuda.is_available():
        return
    (tmp_path / "out_fin_1_of_1").write_text("1 2 0.5\n2 1 0.5\n")
    (tmp_path / "out_test_rat_1_of_1").write_text("1 2147483646 4 \n")
    p = subprocess.run([exe, "100", "1"], cwd=str(tmp_path), capture_output=True, text=True, timeout=60)
    assert p.returncode != 0 and "no CPU fallback" in (p.stdout + p.stderr)
    assert not (tmp_path / "out_res_1_of_1").exists()

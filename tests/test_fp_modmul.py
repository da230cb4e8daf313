"""The FP64 butterflies of the engine (csrc/ntt.cuh FpArith) rely on an error-free FMA modular product; this checks
the same C99 sequence exhaustively-at-random on the CPU against 128-bit integer arithmetic."""
import os
import subprocess


def test_fp64_error_free_modmul(tmp_path):
    src = os.path.join(os.path.dirname(__file__), "fp_modmul_check.c")
    exe = tmp_path / "fpmod"
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-o", str(exe), src, "-lm"])
    out = subprocess.check_output([str(exe)], text=True, timeout=600)
    assert "bad=0" in out, out
    ratio = float(out.split("max|r|/q=")[1])
    assert ratio < 0.63

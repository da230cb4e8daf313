"""The C-ABI library loads on a GPU-less box and exports every symbol include/pfhe_b200.h declares; parameter
validation that needs no device works; compute entry points fail loudly (no CPU fallback)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "pfhe_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pfhe_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import phantom_fhe_b200 as pf
    names = declared_symbols()
    assert len(names) >= 30
    raw = ctypes.CDLL(pf.LIB_PATH)
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/pfhe_b200.h but not exported"
        assert n in pf._lib._sigs or n == "pfhe_last_error", f"{n} has no ctypes signature"


def test_host_only_entry_points():
    import phantom_fhe_b200 as pf
    assert pf.CoeffModulus.Create(4096, [50]) == [1125899906826241]
    p = pf.CoeffModulus.Create(65536, [60, 40, 40, 60])
    assert p == [1152921504598720513, 1099507695617, 1099510054913, 1152921504606584833]  # SURVEY.md probe
    assert pf.get_elt_from_step(1, 65536) == 5 and pf.get_elt_from_step(0, 4096) == 8191
    with pytest.raises(ValueError, match="step count too large"):
        pf.get_elt_from_step(40000, 65536)
    with pytest.raises((ValueError, RuntimeError)):
        pf.CoeffModulus.Create(4096, [70])


def test_no_cpu_fallback():
    import torch
    import phantom_fhe_b200 as pf
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    parms = pf.EncryptionParameters(pf.scheme_type.ckks)
    parms.set_poly_modulus_degree(4096)
    parms.set_coeff_modulus(pf.CoeffModulus.Create(4096, [50, 40, 50]))
    parms.set_special_modulus_size(1)
    with pytest.raises(RuntimeError, match="CUDA"):
        pf.PhantomContext(parms)
    # and straight through the C-ABI
    h = ctypes.c_void_p()
    primes = (ctypes.c_uint64 * 3)(*parms.coeff_modulus)
    rc = pf.lib.pfhe_engine_create(ctypes.byref(h), 3, 4096, primes, 3, 1, 0, None, 0)
    assert rc == 3 and not h.value  # PFHE_ERR_CUDA


def test_product_does_not_touch_the_oracle():
    pkg = os.path.join(ROOT, "phantom-fhe_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "liboracle" not in text and "fhe_oracle" not in text and "harness" not in text, f


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/pfhe_b200.h is a C header (no C++ or torch types in any signature): a C99 translation unit that includes it
    compiles with -Wall -Werror -pedantic, links against the library without any other dependency named on the command
    line, and calls a host-only entry point the way a cgo / JNI / ctypes binding would."""
    import shutil
    import subprocess
    import phantom_fhe_b200 as pf
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no C compiler")
    src = tmp_path / "demo.c"
    src.write_text(
        '#include <stdio.h>\n#include "pfhe_b200.h"\n'
        "int main(void) {\n"
        "    int bits[2] = {50, 40};\n    uint64_t primes[2];\n"
        "    int rc = pfhe_create_primes(4096, bits, 2, primes);\n"
        "    if (rc != 0) { printf(\"error %d: %s\\n\", rc, pfhe_last_error()); return 1; }\n"
        "    printf(\"%llu %llu\\n\", (unsigned long long) primes[0], (unsigned long long) primes[1]);\n"
        "    pfhe_engine *e = 0;\n"
        "    rc = pfhe_engine_create(&e, 3, 4096, primes, 2, 1, 0, 0, 0);   /* no device here: must fail, not fall back */\n"
        "    printf(\"%d\\n\", rc);\n    return 0;\n}\n")
    exe = tmp_path / "demo"
    lib_dir = os.path.dirname(pf.LIB_PATH)
    subprocess.check_call([gcc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src), "-o",
                           str(exe), "-L", lib_dir, "-lpfhe_b200", f"-Wl,-rpath,{lib_dir}"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stdout + out.stderr
    lines = out.stdout.split()
    assert [int(lines[0]), int(lines[1])] == pf.CoeffModulus.Create(4096, [50, 40])
    import torch
    if not torch.cuda.is_available():
        assert int(lines[2]) == 3   # PFHE_ERR_CUDA


def test_cpp_mirror_header_compiles():
    """include/phantom_b200.hpp (the reference's C++ class and function names over the C-ABI) and the demo application
    written against it compile as C++17 with warnings as errors; without a device the application fails loudly."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    cuda_inc = "/usr/local/cuda/include"
    if gxx is None or not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("no C++ compiler or CUDA headers")
    subprocess.check_call([gxx, "-std=c++17", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-I", os.path.join(ROOT, "include"),
                           "-I", cuda_inc, os.path.join(ROOT, "tests", "cpp", "mirror_demo.cpp")])
    exe = os.path.join(ROOT, "tests", "cpp", "mirror_demo")
    import torch
    if os.path.exists(exe) and not torch.cuda.is_available():
        out = subprocess.run([exe], capture_output=True, text=True, timeout=120)
        assert out.returncode == 2 and "no CPU path" in out.stdout

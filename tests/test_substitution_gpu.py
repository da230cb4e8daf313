"""INTEGRATION.md section 3 compiled and run: oracle/_ref/libphantom_subst.so is the reference's UNMODIFIED evaluate.cu /
rns.cu / secretkey.cu / ciphertext.h linked against oracle/subst/phantom_on_pfhe.cu, which defines the reference's
kernel-level launchers (nwt_2d_radix8_*, DRNSTool::modup, DRNSTool::moddown_from_NTT, key_switch_inner_prod) on top of
libpfhe_b200.so.  The same words and keys go through the stock library and through the substituted one; and the
reference's own example program (examples/example.cu: `example_context 3`, 2, 1) runs its checks on the new kernels."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

import harness as H
from harness import P

pytestmark = pytest.mark.gpu

REF_DIR = os.path.join(H.ORACLE_DIR, "_ref")
SUBST_SO = os.path.join(REF_DIR, "libphantom_subst.so")
HERE = os.path.dirname(os.path.abspath(__file__))


def _need(path):
    if not os.path.exists(path):
        pytest.skip(f"{os.path.relpath(path, H.ROOT)} was not built (oracle/Makefile.subst)")


@pytest.mark.parametrize("name", ["ckks_primary", "ckks_alpha1", "bgv", "bfv_hps"])
def test_reference_evaluator_on_the_substituted_kernels(tmp_path, name):
    r = H.reference()
    if r is None:
        pytest.skip("oracle/_ref/libphantom_ref.so was not built")
    _need(SUBST_SO)
    ps, mul_tech = {
        "ckks_primary": (H.params_primary(), 0),
        "ckks_alpha1": (H.params_secondary(), 0),
        "bgv": (H.params_small(16384, l=6, alpha=2, scheme=1, t=65537), 0),
        "bfv_hps": (H.params_bfv_bench(2), 2),
    }[name]
    steps = (ctypes.c_int * 1)(1)
    h = r.ref_create(ps.scheme, ps.n, P(ps.primes), ps.size_QP, ps.size_P, ps.t, mul_tech, steps, 1, float(2 ** 40), 1)
    assert h, r.ref_last_error()
    try:
        dnum, l, n = r.ref_dnum(h), ps.limbs(), ps.n
        rlk = np.zeros((dnum, 2, ps.size_QP, n), dtype=np.uint64)
        glk = np.zeros_like(rlk)
        for d in range(dnum):
            assert r.ref_key_get(h, -1, d, P(rlk[d])) == 0
            assert r.ref_key_get(h, 0, d, P(glk[d])) == 0
        a, b = H.ciphertext(ps, 1), H.ciphertext(ps, 2)
        want_mul = np.zeros((2, l, n), dtype=np.uint64)
        assert r.ref_multiply_relin(h, 1, P(a), P(b), P(want_mul)) == 0, r.ref_last_error()
        want_rot = np.zeros((2, l, n), dtype=np.uint64)
        assert r.ref_rotate(h, 1, P(a), 1, P(want_rot)) == 0, r.ref_last_error()
        want_down = np.zeros((2, l - 1, n), dtype=np.uint64)
        if ps.scheme == 3:
            assert r.ref_rescale(h, 1, P(a), 2, P(want_down)) == 0, r.ref_last_error()
        else:
            assert r.ref_mod_switch(h, 1, P(a), 2, P(want_down)) == 0, r.ref_last_error()
        times = (ctypes.c_double * 60)()
        assert r.ref_time_op(h, 0, 1, P(a), P(b), 0, 0, 60, times) == 0
        ref_us = sorted(times[10:])[25]
    finally:
        r.ref_destroy(h)
    w = str(tmp_path)
    np.save(f"{w}/meta.npy", np.array([ps.scheme, n, ps.size_P, ps.t, mul_tech], dtype=np.int64))
    np.save(f"{w}/primes.npy", ps.primes)
    np.save(f"{w}/rlk.npy", rlk), np.save(f"{w}/glk.npy", glk), np.save(f"{w}/a.npy", a), np.save(f"{w}/b.npy", b)
    env = dict(os.environ, PFHE_REF_SO=SUBST_SO, PYTHONPATH=HERE + os.pathsep + H.ROOT)
    p = subprocess.run([sys.executable, os.path.join(HERE, "subst_worker.py"), w], env=env, capture_output=True, text=True,
                       timeout=900)
    assert p.returncode == 0, p.stdout[-2000:] + p.stderr[-4000:]
    assert np.array_equal(np.load(f"{w}/got_mul.npy"), want_mul), "multiply_inplace + relinearize_inplace"
    assert np.array_equal(np.load(f"{w}/got_rot.npy"), want_rot), "rotate_inplace"
    assert np.array_equal(np.load(f"{w}/got_down.npy"), want_down), "rescale_to_next / mod_switch_to_next"
    sub_us = np.load(f"{w}/time_us.npy")[25]
    print(f"\n[{name}] reference evaluate.cu HMult+Relin: stock kernels {ref_us:.1f} us, substituted kernels {sub_us:.1f} us")


@pytest.mark.parametrize("selection", ["3", "2", "1"])
def test_reference_examples_run_on_the_substituted_kernels(selection):
    """examples/example.cu: 3 = CKKS (encode / encrypt / add / HomMul / rotate checks over alpha in {1, 2, 3, 4, 15}),
    2 = BGV, 1 = BFV; the program throws when one of its own checks fails"""
    exe = os.path.join(REF_DIR, "example_context_subst")
    _need(exe)
    p = subprocess.run([exe, selection], capture_output=True, text=True, timeout=1500)
    tail = p.stdout[-1500:] + p.stderr[-1500:]
    assert p.returncode == 0, tail
    assert "error" not in p.stderr.lower(), tail

"""Host-side logic of the mirror interface that needs no GPU: NAF rotation decomposition, parameter plumbing,
batch sharding (incl. a world_size-2 gloo run)."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_naf_matches_reference_definition():
    from phantom_fhe_b200.api import _naf
    for v in list(range(-70, 71)) + [1000, -1234, 32767]:
        parts = _naf(v)
        assert sum(parts) == v
        mags = sorted(abs(p) for p in parts)
        assert all(m & (m - 1) == 0 for m in mags)            # powers of two
        assert all(b >= 4 * a or b >= 2 * a for a, b in zip(mags, mags[1:]))
        for a, b in zip(mags, mags[1:]):
            assert b != 2 * a                                   # non-adjacent
    assert _naf(3) == [-1, 4] and _naf(7) == [-1, 8] and _naf(0) == []


def test_encryption_parameters_errors():
    import phantom_fhe_b200 as pf
    p = pf.EncryptionParameters()
    with pytest.raises(RuntimeError):
        p.set_coeff_modulus([97])
    p = pf.EncryptionParameters(pf.scheme_type.ckks)
    with pytest.raises(RuntimeError):
        p.set_plain_modulus(65537)
    p.set_special_modulus_size(4)
    assert p.special_modulus_size == 4


def test_shard_range_partitions_everything():
    from phantom_fhe_b200.shard import shard_range
    for total in (0, 1, 7, 8, 1024, 1000):
        for world in (1, 2, 3, 4, 8):
            got = []
            for r in range(world):
                b, e = shard_range(total, r, world)
                assert 0 <= b <= e <= total
                got.extend(range(b, e))
            assert got == list(range(total))
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


WORKER = r"""
import os, sys
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
from phantom_fhe_b200.shard import shard_range, max_over_ranks
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:{port}", rank=int(sys.argv[1]), world_size=2)
rank = dist.get_rank()
b, e = shard_range(1024, rank, 2)
mine = torch.zeros(1024, dtype=torch.int64); mine[b:e] = 1
dist.all_reduce(mine)
assert int(mine.sum()) == 1024 and int(mine.max()) == 1          # disjoint cover, no collective on the data path
t = max_over_ranks(10.0 + rank, dist)
assert t == 11.0
dist.barrier(); dist.destroy_process_group()
print("ok", rank)
"""


def test_two_rank_gloo_sharding(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT, port=29000 + os.getpid() % 2000))
    procs = [subprocess.Popen([sys.executable, str(script), str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                              text=True) for r in range(2)]
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o


def test_serialisation_round_trip_and_layout():
    """phantom-fhe_b200/serial.py: the reference's stream layouts (include/ciphertext.h:173-213, secretkey.h:129-219,346-390)
    -- header bytes at their offsets, round trips of ciphertexts, relinearisation / Galois keys and secret keys, truncated
    streams refused.  (Byte equality with streams written by the unmodified reference: tests/test_gpu_parity.py.)"""
    import importlib.util
    import io
    import struct
    spec = importlib.util.spec_from_file_location("pfhe_serial", os.path.join(ROOT, "phantom-fhe_b200", "serial.py"))
    serial = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(serial)
    rng = np.random.default_rng(0)
    words = rng.integers(0, 1 << 60, (3, 2, 16), dtype=np.uint64)
    buf = io.BytesIO()
    serial.write_ciphertext(buf, words, 2, scale=2.0 ** 40, correction_factor=7, noise_scale_deg=3, is_ntt_form=True,
                            is_asymmetric=True)
    raw = buf.getvalue()
    assert len(raw) == 58 + words.size * 8
    assert struct.unpack_from("<QQQQ", raw, 0) == (2, 3, 16, 2)
    assert struct.unpack_from("<d", raw, 32)[0] == 2.0 ** 40 and struct.unpack_from("<QQ", raw, 40) == (7, 3)
    assert raw[56:58] == b"\x01\x01"
    back, hdr = serial.read_ciphertext(io.BytesIO(raw))
    assert np.array_equal(back, words) and hdr["chain_index"] == 2 and hdr["noise_scale_deg"] == 3 and hdr["is_asymmetric"]
    with pytest.raises(ValueError):
        serial.read_ciphertext(io.BytesIO(raw[:-8]))
    keys = [[rng.integers(0, 1 << 50, (2, 3, 16), dtype=np.uint64) for _ in range(2)] for _ in range(3)]
    buf = io.BytesIO()
    serial.write_galois_key(buf, keys)
    got = serial.read_galois_key(io.BytesIO(buf.getvalue()))
    assert len(got) == 3 and all(np.array_equal(a, b) for ka, kb in zip(got, keys) for a, b in zip(ka, kb))
    assert struct.unpack_from("<QQ", buf.getvalue(), 0) == (3, 2)   # key count, then dnum of the first key
    pw = rng.integers(0, 1 << 50, (2, 3, 16), dtype=np.uint64)
    buf = io.BytesIO()
    serial.write_secret_key(buf, pw)
    assert struct.unpack_from("<QQQ", buf.getvalue(), 0) == (2, 16, 3)
    assert np.array_equal(serial.read_secret_key(io.BytesIO(buf.getvalue())), pw)
    # seed-compressed symmetric form (save_symmetric / load_symmetric, ciphertext.h:216-307): header, c0, 64-byte seed
    c0, seed = words[0], bytes(range(64))
    buf = io.BytesIO()
    serial.write_ciphertext_symmetric(buf, c0, seed, 1, scale=2.0 ** 30, is_ntt_form=False)
    raw = buf.getvalue()
    assert len(raw) == 58 + c0.size * 8 + 64 and raw[-64:] == seed and raw[56:58] == b"\x00\x00"
    assert struct.unpack_from("<QQQQ", raw, 0) == (1, 2, 16, 2)
    back, sd, hdr = serial.read_ciphertext_symmetric(io.BytesIO(raw))
    assert np.array_equal(back, c0) and sd == seed and hdr["scale"] == 2.0 ** 30 and not hdr["is_ntt_form"]
    with pytest.raises(ValueError):
        serial.read_ciphertext_symmetric(io.BytesIO(raw[:-1]))
    with pytest.raises(ValueError):
        serial.write_ciphertext_symmetric(io.BytesIO(), c0, b"short", 1)
    full = io.BytesIO()
    serial.write_ciphertext(full, words, 1, is_asymmetric=True)
    with pytest.raises(RuntimeError):   # "Asymmetric ciphertext does not have seed."
        serial.read_ciphertext_symmetric(io.BytesIO(full.getvalue()))


def test_balance_correction_factors():
    """balance_correction_factors (src/evaluate.cu:14-72) in the host mirror: e1 * f1 = e2 * f2 = f mod t with e1, e2
    invertible, never worse than the trivial pair (f2 / f1, 1), and the small cases worked by hand."""
    import importlib.util
    import math
    import types
    import sys
    src = open(os.path.join(ROOT, "phantom-fhe_b200", "api.py")).read()
    start = src.index("def balance_correction_factors")
    end = src.index("def _check_pair")
    ns = {"math": math}
    exec(src[start:end], ns)   # the function is pure Python; api.py itself needs the CUDA library to import
    bal = ns["balance_correction_factors"]
    t = 65537
    assert bal(3, 5, t) == (15, 5, 3)
    assert bal(1, 1, t) == (1, 1, 1)
    rng = np.random.default_rng(0)
    for _ in range(200):
        f1, f2 = int(rng.integers(1, t)), int(rng.integers(1, t))
        f, e1, e2 = bal(f1, f2, t)
        assert e1 * f1 % t == f and e2 * f2 % t == f and math.gcd(e1, t) == 1 and math.gcd(e2, t) == 1

        def mag(x):
            return min(x, t - x)
        ratio = pow(f1, -1, t) * f2 % t
        assert mag(e1) + mag(e2) <= mag(ratio) + 1
    with pytest.raises(RuntimeError):
        bal(0, 5, t)


def test_plaintext_and_public_key_streams():
    """PhantomPlaintext::save / load (include/plaintext.h:69-97) and PhantomPublicKey::save / load (include/secretkey.h:85-96):
    field order and sizes, round trips, truncation refused."""
    import importlib.util
    import io
    import struct
    spec = importlib.util.spec_from_file_location("pfhe_serial2", os.path.join(ROOT, "phantom-fhe_b200", "serial.py"))
    serial = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(serial)
    rng = np.random.default_rng(1)
    words = rng.integers(0, 1 << 50, (3, 16), dtype=np.uint64)
    buf = io.BytesIO()
    serial.write_plaintext(buf, words, 2, scale=2.0 ** 40)
    raw = buf.getvalue()
    assert len(raw) == 32 + words.size * 8 and struct.unpack_from("<QQQd", raw, 0) == (2, 16, 3, 2.0 ** 40)
    back, ci, scale = serial.read_plaintext(io.BytesIO(raw))
    assert np.array_equal(back, words) and ci == 2 and scale == 2.0 ** 40
    buf = io.BytesIO()
    serial.write_plaintext(buf, words[0], 0)   # BFV / BGV: one row of residues mod t
    assert struct.unpack_from("<QQQd", buf.getvalue(), 0) == (0, 16, 1, 1.0)
    with pytest.raises(ValueError):
        serial.read_plaintext(io.BytesIO(raw[:-1]))
    pk = rng.integers(0, 1 << 50, (2, 4, 16), dtype=np.uint64)
    buf = io.BytesIO()
    serial.write_public_key(buf, pk)
    raw = buf.getvalue()
    assert struct.unpack_from("<QQQQ", raw, 0) == (0, 2, 16, 4) and raw[56:58] == b"\x01\x00"
    assert np.array_equal(serial.read_public_key(io.BytesIO(raw)), pk)
    one = io.BytesIO()
    serial.write_ciphertext(one, pk, 1)
    with pytest.raises(ValueError):
        serial.read_public_key(io.BytesIO(one.getvalue()))


def test_cpp_mirror_balances_like_the_python_mirror(tmp_path):
    """include/phantom_b200.hpp's balance_correction_factors against the Python mirror's (both restate
    src/evaluate.cu:14-72) on a few hundred factor pairs, several plain moduli; and its non-adjacent form of rotation steps
    against the Python mirror's."""
    import math
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    cuda_inc = "/usr/local/cuda/include"
    if gxx is None or not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("no C++ compiler or CUDA headers")
    src = tmp_path / "bal.cpp"
    src.write_text(
        '#include <cstdio>\n#include <cstdlib>\n#include "phantom_b200.hpp"\n'
        "int main(int argc, char **argv) {\n"
        "    for (int i = 1; i + 2 < argc; i += 3) {\n"
        "        auto b = phantom_b200::detail::balance_correction_factors(strtoull(argv[i], 0, 10), strtoull(argv[i + 1], 0, 10),\n"
        "                                                                  strtoull(argv[i + 2], 0, 10));\n"
        '        std::printf("%llu %llu %llu\\n", (unsigned long long) b.f, (unsigned long long) b.e1, (unsigned long long) b.e2);\n'
        "    }\n"
        "    for (int step = -70; step <= 70; step++) {\n"
        '        std::printf("naf %d:", step);\n'
        '        for (int p : phantom_b200::detail::naf(step)) std::printf(" %d", p);\n'
        '        std::printf("\\n");\n'
        "    }\n    return 0;\n}\n")
    exe = tmp_path / "bal"
    subprocess.check_call([gxx, "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), "-I", cuda_inc, str(src), "-o", str(exe)])
    api_src = open(os.path.join(ROOT, "phantom-fhe_b200", "api.py")).read()
    ns = {"math": math}
    exec(api_src[api_src.index("def balance_correction_factors"):api_src.index("def _check_pair")], ns)
    rng = np.random.default_rng(5)
    cases = []
    for t in (65537, 1032193, 786433, (1 << 20) + 7):
        for _ in range(60):
            f1, f2 = int(rng.integers(1, t)), int(rng.integers(1, t))
            if math.gcd(f1, t) == 1:
                cases.append((f1, f2, t))
    args = [str(v) for c in cases for v in c]
    out = subprocess.run([str(exe)] + args, capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    lines = out.stdout.strip().splitlines()
    got = [tuple(int(v) for v in line.split()) for line in lines if not line.startswith("naf")]
    assert got == [ns["balance_correction_factors"](*c) for c in cases]
    exec(api_src[api_src.index("def _naf"):api_src.index("def rotate_inplace")], ns)
    nafs = {int(line.split(":")[0].split()[1]): [int(v) for v in line.split(":")[1].split()] for line in lines if line.startswith("naf")}
    assert len(nafs) == 141 and all(nafs[step] == ns["_naf"](step) for step in range(-70, 71))


def test_pyphantom_module_has_the_binding_surface():
    """`import pyPhantom` offers every class, enum and function name the reference's pybind module defines
    (python/src/binding.cu:13-166) and the methods the binding attaches to them."""
    import pyPhantom as phantom
    for name in ("scheme_type", "mul_tech_type", "sec_level_type", "modulus", "create_coeff_modulus", "create_plain_modulus",
                 "params", "cuda_stream", "context", "secret_key", "public_key", "relin_key", "galois_key", "get_elt_from_step",
                 "get_elts_from_steps", "batch_encoder", "ckks_encoder", "plaintext", "ciphertext", "negate", "add", "add_plain",
                 "add_many", "sub", "sub_plain", "multiply", "multiply_and_relin", "multiply_plain", "relinearize",
                 "rescale_to_next", "mod_switch_to_next", "mod_switch_to", "apply_galois", "rotate", "hoisting"):
        assert hasattr(phantom, name), name
    for cls, methods in ((phantom.params, ("set_mul_tech", "set_poly_modulus_degree", "set_special_modulus_size", "set_galois_elts",
                                           "set_coeff_modulus", "set_plain_modulus")),
                         (phantom.secret_key, ("gen_publickey", "gen_relinkey", "create_galois_keys", "encrypt_symmetric", "decrypt")),
                         (phantom.public_key, ("encrypt_asymmetric",)),
                         (phantom.batch_encoder, ("slot_count", "encode", "decode")),
                         (phantom.ckks_encoder, ("slot_count", "encode_complex_vector", "encode_double_vector",
                                                 "decode_complex_vector", "decode_double_vector")),
                         (phantom.ciphertext, ("set_scale",))):
        for m in methods:
            assert callable(getattr(cls, m, None)), f"{cls.__name__}.{m}"
    assert [int(v) for v in (phantom.scheme_type.bgv, phantom.scheme_type.bfv, phantom.scheme_type.ckks)] == [1, 2, 3]
    assert int(phantom.mul_tech_type.hps_overq_leveled) == 4
    assert phantom.create_plain_modulus(8192, 20) % 16384 == 1
    p = phantom.plaintext()
    assert p.chain_index() == 0 and p.scale() == 1.0 and phantom.ciphertext().size() == 0


class _StubLib:
    """stands in for libpfhe_b200 in host-logic tests: every entry point succeeds and records its name"""

    def __init__(self):
        self.calls = []

    def __getattr__(self, name):
        def call(*args):
            self.calls.append(name)
            if name == "pfhe_find_levels_to_drop":
                args[-1]._obj.value = 1
            return 0
        return call


def test_evaluator_bookkeeping_with_a_stub_library(monkeypatch):
    """The host side of the mirror's evaluator without a device: which C-ABI entry point each call reaches, and the
    bookkeeping the reference does around them -- CKKS scales multiply, BGV correction factors multiply (products) or are
    balanced (sums) or pick up q_last^-1 (mod switch), hps_overq_leveled raises the noise degree, the reference's operand
    checks fire with its messages (src/evaluate.cu:115-197, 345-397, 1029-1104, 1376-1427, 1505-1543)."""
    import torch
    import phantom_fhe_b200.api as api
    stub = _StubLib()
    monkeypatch.setattr(api, "lib", stub)
    monkeypatch.setattr(api, "_stream", lambda: None)

    class Ctx:
        def __init__(self, scheme, mul_tech=api.mul_tech_type.hps, t=65537):
            self.scheme, self._h, self.poly_degree, self.size_Q, self.size_QP, self.size_P = scheme, None, 8, 3, 4, 1
            self.parms = types.SimpleNamespace(mul_tech=mul_tech, plain_modulus=t, coeff_modulus=[97, 193, 257, 769], galois_elts=[])
            self.device = torch.device("cpu")

        def coeff_modulus_size(self, chain_index):
            return self.size_Q - (chain_index - 1)

    import types

    def ct(ctx, size=2, chain_index=1, scale=1.0, cf=1):
        c = api.PhantomCiphertext(None, torch.zeros((size, ctx.coeff_modulus_size(chain_index), 8), dtype=torch.int64), chain_index,
                                  scale, ctx.scheme != api.scheme_type.bfv)
        c.correction_factor = cf
        return c

    ckks = Ctx(api.scheme_type.ckks)
    a, b = ct(ckks, scale=2.0 ** 20), ct(ckks, scale=2.0 ** 20)
    api.multiply_inplace(ckks, a, b)
    assert stub.calls[-1] == "pfhe_multiply" and a.size() == 3 and a.scale == 2.0 ** 40
    api.relinearize_inplace(ckks, a, types.SimpleNamespace(public_keys_ptr=lambda: None))
    assert stub.calls[-1] == "pfhe_relinearize_inplace" and a.size() == 2
    with pytest.raises(ValueError, match="scale mismatch"):
        api.multiply_inplace(ckks, a, b)
    with pytest.raises(ValueError, match="parameter mismatch"):
        api.multiply_and_relin_inplace(ckks, ct(ckks), ct(ckks, chain_index=2), None)
    nxt = api.rescale_to_next(ckks, b)
    assert nxt.chain_index == 2 and nxt.scale == 2.0 ** 20 / 257 and nxt.coeff_modulus_size() == 2
    api.rescale_to_next_inplace(ckks, b)
    assert b.chain_index == 2 and b.coeff_modulus_size() == 2
    api.mod_switch_to_next_inplace(ckks, b)
    assert b.chain_index == 3 and stub.calls[-1] == "pfhe_mod_switch_to_next"
    api.keyswitch_inplace(ckks, b, torch.zeros((1, 8), dtype=torch.int64), types.SimpleNamespace(public_keys_ptr=lambda: None))
    assert stub.calls[-1] == "pfhe_keyswitch_inplace"

    bgv = Ctx(api.scheme_type.bgv)
    a, b = ct(bgv, cf=3), ct(bgv, cf=5)
    api.multiply_and_relin_inplace(bgv, a, b, types.SimpleNamespace(public_keys_ptr=lambda: None))
    assert stub.calls[-1] == "pfhe_multiply_and_relin" and a.correction_factor == 15
    low = api.mod_switch_to_next(bgv, a)
    assert low.chain_index == 2 and low.correction_factor == 15 * pow(257, -1, 65537) % 65537
    a, b = ct(bgv, cf=3), ct(bgv, cf=5)
    api.add_inplace(bgv, a, b)
    assert stub.calls.count("pfhe_multiply_scalar_rns_poly") == 2 and a.correction_factor == api.balance_correction_factors(3, 5, 65537)[0]
    assert b.correction_factor == 5
    api.sub_inplace(bgv, a, ct(bgv, cf=a.correction_factor), negate=True)
    assert stub.calls[-1] == "pfhe_sub_rns_poly"
    api.add_plain_inplace(bgv, a, torch.zeros(8, dtype=torch.int64))
    assert stub.calls[-1] == "pfhe_add_plain_inplace"
    with pytest.raises(ValueError, match="poly number mismatch"):
        api.add_inplace(bgv, a, ct(bgv, size=3))

    lev = Ctx(api.scheme_type.bfv, api.mul_tech_type.hps_overq_leveled)
    a, b = ct(lev), ct(lev)
    a.is_asymmetric = True
    api.multiply_and_relin_inplace(lev, a, b, types.SimpleNamespace(public_keys_ptr=lambda: None))
    assert stub.calls[-1] == "pfhe_multiply_and_relin_leveled" and a.noise_scale_deg == 2
    api.multiply_inplace(lev, a, b)
    assert stub.calls[-1] == "pfhe_multiply_leveled" and a.noise_scale_deg == 3 and a.size() == 3
    low = api.mod_switch_to_next(lev, b)
    assert low.noise_scale_deg == 1 and not low.is_asymmetric and not low.is_ntt_form
    api.keyswitch_inplace(lev, ct(lev), torch.zeros((3, 8), dtype=torch.int64), types.SimpleNamespace(public_keys_ptr=lambda: None), False)
    assert stub.calls[-1] == "pfhe_keyswitch_leveled_inplace"
    with pytest.raises(ValueError, match="BFV encrypted cannot be in NTT form"):
        wrong = ct(lev)
        wrong.is_ntt_form = True
        api.multiply_plain_inplace(lev, wrong, torch.zeros(8, dtype=torch.int64))
    total = api.add_many(lev, [ct(lev), ct(lev), ct(lev)])
    assert total.size() == 2 and stub.calls[-1] == "pfhe_add_rns_poly"
    # BFV hoisting under hps_overq_leveled: one call of the leveled entry point with the levels of a key switch
    lev.parms.galois_elts = [0]   # the stub's get_elt_from_step answers 0 for every step
    glk = types.SimpleNamespace(get_relin_keys=lambda idx: types.SimpleNamespace(public_keys_ptr=lambda: ctypes.c_void_p(0)))
    c = ct(lev)
    before = len(stub.calls)
    api.hoisting_inplace(lev, c, glk, [1, 2])
    made = stub.calls[before:]
    assert made[-1] == "pfhe_hoisting_leveled_inplace" and "pfhe_find_levels_to_drop" in made
    assert "pfhe_hoisting_inplace" not in made and c.size() == 2
    low2 = ct(lev)
    low2.chain_index = 2
    with pytest.raises(ValueError, match="first data level"):
        api.hoisting_inplace(lev, low2, glk, [1])


def test_pyphantom_wrappers_with_a_stub_library(monkeypatch):
    """The binding-shaped layer (phantom-fhe_b200/pyphantom.py) without a device: results come back as `ciphertext`
    objects, operands are left alone, plaintext levels are checked and switched like the reference's overloads
    (include/evaluate.cuh:150-207)."""
    import types
    import torch
    import phantom_fhe_b200.api as api
    import pyPhantom as phantom
    stub = _StubLib()
    monkeypatch.setattr(api, "lib", stub)
    monkeypatch.setattr(api, "_stream", lambda: None)
    ctx = types.SimpleNamespace(scheme=api.scheme_type.ckks, _h=None, poly_degree=8, size_Q=3, size_QP=4, size_P=1,
                                device=torch.device("cpu"), coeff_modulus_size=lambda ci: 3 - (ci - 1),
                                parms=types.SimpleNamespace(mul_tech=api.mul_tech_type.none, plain_modulus=0,
                                                            coeff_modulus=[97, 193, 257, 769], galois_elts=[]))
    a = phantom.ciphertext(None, torch.zeros((2, 3, 8), dtype=torch.int64), 1, 2.0 ** 20, True)
    b = phantom.ciphertext(None, torch.ones((2, 3, 8), dtype=torch.int64), 1, 2.0 ** 20, True)
    keys = types.SimpleNamespace(public_keys_ptr=lambda: None)
    prod = phantom.multiply_and_relin(ctx, a, b, keys)
    assert isinstance(prod, phantom.ciphertext) and prod.scale == 2.0 ** 40 and a.scale == 2.0 ** 20 and prod is not a
    total = phantom.add(ctx, a, b)
    assert isinstance(total, phantom.ciphertext) and stub.calls[-1] == "pfhe_add_rns_poly"
    dest = phantom.ciphertext()
    assert phantom.add_many(ctx, [a, b], dest) is dest and dest.size() == 2
    pt = phantom.plaintext(torch.zeros((3, 8), dtype=torch.int64), 1, 2.0 ** 20)
    assert phantom.add_plain(ctx, a, pt).chain_index == 1
    low = phantom.mod_switch_to(ctx, pt, 3)
    assert low.chain_index() == 3 and low.data.shape[0] == 1 and pt.chain_index() == 1
    with pytest.raises(ValueError, match="parameter mismatch"):
        phantom.add_plain(ctx, a, low)
    with pytest.raises(ValueError, match="higher level"):
        phantom.mod_switch_to(ctx, low, 1)
    with pytest.raises(ValueError, match="end of modulus switching chain"):
        phantom.mod_switch_to_next(ctx, low)
    down = phantom.mod_switch_to(ctx, a, 2)
    assert isinstance(down, phantom.ciphertext) and down.chain_index == 2 and a.chain_index == 1
    scaled = phantom.multiply_plain(ctx, a, pt)
    assert scaled.scale == 2.0 ** 40


def test_cpp_mirror_application_on_the_oracle(tmp_path):
    """tests/cpp/mirror_demo.cpp (an application written against the reference's C++ names, include/phantom_b200.hpp) linked
    against tests/cpp/mock_pfhe_oracle.cpp -- host memory for device memory, the CPU oracle behind every C-ABI entry point
    the mirror calls -- instead of the real library: the mirror's call sequences, buffer sizes and bookkeeping carry BFV
    (hps_overq_leveled), BGV and CKKS flows from key generation to decoded results.  (The GPU suite runs the same
    application on the real library.)"""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    cuda_inc = "/usr/local/cuda/include"
    if gxx is None or not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("no C++ compiler or CUDA headers")
    oracle_dir = os.path.join(ROOT, "oracle")
    exe = tmp_path / "mirror_demo_cpu"
    subprocess.check_call([gxx, "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), "-I", cuda_inc,
                           os.path.join(ROOT, "tests", "cpp", "mirror_demo.cpp"), os.path.join(ROOT, "tests", "cpp", "mock_pfhe_oracle.cpp"),
                           "-o", str(exe), "-L", oracle_dir, "-loracle", f"-Wl,-rpath,{oracle_dir}"])
    dump = tmp_path / "streams"
    dump.mkdir()
    for logn in ("12", "13"):
        out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600,
                             env=dict(os.environ, PFHE_DEMO_LOGN=logn, PFHE_DEMO_DUMP=str(dump)))
        assert out.returncode == 0 and out.stdout.strip().endswith("OK"), out.stdout + out.stderr
        assert out.stdout.count("ok  ") == 53 and "FAIL" not in out.stdout
    for drop in ("1", "2"):   # the level-dropping branches of hps_overq_leveled: products, relinearisation, rotations, hoisting
        out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=600,
                             env=dict(os.environ, PFHE_DEMO_LOGN="12", PFHE_MOCK_DROP=drop))
        assert out.returncode == 0 and out.stdout.count("ok  ") == 53 and "FAIL" not in out.stdout, out.stdout + out.stderr
    # the streams the C++ mirror wrote are the Python mirror's (and so the reference's) formats: read and re-written byte for byte
    import importlib.util
    import io
    spec = importlib.util.spec_from_file_location("pfhe_serial3", os.path.join(ROOT, "phantom-fhe_b200", "serial.py"))
    serial = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(serial)
    n = 8192
    for scheme in ("bfv", "bgv"):
        def raw(kind):
            return open(dump / f"{scheme}_{kind}.bin", "rb").read()

        def rewritten(write, *args, **kw):
            buf = io.BytesIO()
            write(buf, *args, **kw)
            return buf.getvalue()
        words, hdr = serial.read_ciphertext(io.BytesIO(raw("ciphertext")))
        assert words.shape == (2, 4, n) and hdr["is_asymmetric"] and hdr["is_ntt_form"] == (scheme == "bgv")
        assert rewritten(serial.write_ciphertext, words, hdr["chain_index"], hdr["scale"], hdr["correction_factor"],
                         hdr["noise_scale_deg"], hdr["is_ntt_form"], hdr["is_asymmetric"]) == raw("ciphertext")
        c0, seed, hdr = serial.read_ciphertext_symmetric(io.BytesIO(raw("ciphertext_symmetric")))
        assert rewritten(serial.write_ciphertext_symmetric, c0, seed, hdr["chain_index"], hdr["scale"], hdr["correction_factor"],
                         hdr["noise_scale_deg"], hdr["is_ntt_form"]) == raw("ciphertext_symmetric")
        pk = serial.read_public_key(io.BytesIO(raw("public_key")))
        assert pk.shape == (2, 6, n) and rewritten(serial.write_public_key, pk) == raw("public_key")
        digits = serial.read_relin_key(io.BytesIO(raw("relin_key")))
        assert len(digits) == 2 and rewritten(serial.write_relin_key, digits) == raw("relin_key")
        keys = serial.read_galois_key(io.BytesIO(raw("galois_key")))
        assert len(keys) == 1 and rewritten(serial.write_galois_key, keys) == raw("galois_key")
        powers = serial.read_secret_key(io.BytesIO(raw("secret_key")))
        assert powers.shape == (2, 6, n) and rewritten(serial.write_secret_key, powers) == raw("secret_key")
        plain, ci, scale = serial.read_plaintext(io.BytesIO(raw("plaintext")))
        assert plain.shape == (1, n) and ci == 0 and rewritten(serial.write_plaintext, plain, ci, scale) == raw("plaintext")
    plain, ci, scale = serial.read_plaintext(io.BytesIO(open(dump / "ckks_plaintext.bin", "rb").read()))
    assert plain.shape == (3, n) and ci == 1 and scale == 2.0 ** 40


def test_cpp_mirror_application_under_sanitizers(tmp_path):
    """The same application, stand-in and the oracle's C source compiled with AddressSanitizer + UndefinedBehaviorSanitizer:
    "device" buffers are heap blocks here, so a ciphertext, key or plaintext buffer sized wrongly by the mirror (or overrun
    by the oracle) is an error, and so is any leak of the RAII wrappers."""
    import shutil
    import subprocess
    gcc, gxx = shutil.which("gcc"), shutil.which("g++")
    cuda_inc = "/usr/local/cuda/include"
    if gcc is None or gxx is None or not os.path.exists(os.path.join(cuda_inc, "cuda_runtime.h")):
        pytest.skip("no C/C++ compiler or CUDA headers")
    flags = ["-O1", "-g", "-fsanitize=address,undefined", "-fno-omit-frame-pointer"]
    probe = subprocess.run([gxx, "-fsanitize=address,undefined", "-x", "c++", "-", "-o", str(tmp_path / "probe")],
                           input="int main(){return 0;}", capture_output=True, text=True)
    if probe.returncode != 0:
        pytest.skip("sanitizer runtimes are not installed")
    obj = tmp_path / "oracle_asan.o"
    subprocess.check_call([gcc] + flags + ["-w", "-c", os.path.join(ROOT, "oracle", "fhe_oracle.c"), "-I", os.path.join(ROOT, "oracle"),
                                           "-o", str(obj)])
    exe = tmp_path / "mirror_demo_asan"
    subprocess.check_call([gxx, "-std=c++17"] + flags + ["-I", os.path.join(ROOT, "include"), "-I", cuda_inc,
                                                       os.path.join(ROOT, "tests", "cpp", "mirror_demo.cpp"),
                                                       os.path.join(ROOT, "tests", "cpp", "mock_pfhe_oracle.cpp"), str(obj), "-o", str(exe), "-lm"])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=900,
                         env=dict(os.environ, PFHE_DEMO_LOGN="12", ASAN_OPTIONS="detect_leaks=1", UBSAN_OPTIONS="halt_on_error=1"))
    assert out.returncode == 0 and out.stdout.strip().endswith("OK") and "runtime error" not in out.stderr, out.stdout[-2000:] + out.stderr[-4000:]

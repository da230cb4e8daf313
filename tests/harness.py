"""Test / bench harness: ctypes views of the CPU oracle (oracle/liboracle.so) and -- when it was built in the
container that has /root/reference -- of the unmodified reference (oracle/_ref/libphantom_ref.so), plus the
synthetic-input generators of SURVEY.md section 8d.  Test infrastructure only: nothing in phantom-fhe_b200/ imports
this module."""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "liboracle.so")
REF_SO = os.environ.get("PFHE_REF_SO", os.path.join(ORACLE_DIR, "_ref", "libphantom_ref.so"))

u64p = ctypes.POINTER(ctypes.c_uint64)
u32p = ctypes.POINTER(ctypes.c_uint32)
i32p = ctypes.POINTER(ctypes.c_int)
vp = ctypes.c_void_p


def P(a):
    return a.ctypes.data_as(u64p)


def build_oracle():
    src = os.path.join(ORACLE_DIR, "fhe_oracle.c")
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
    return ORACLE_SO


_oracle = None


def oracle():
    global _oracle
    if _oracle is None:
        o = ctypes.CDLL(build_oracle())
        o.orc_create.restype = vp
        o.orc_create.argtypes = [ctypes.c_int, ctypes.c_uint64, u64p, ctypes.c_int, ctypes.c_int, ctypes.c_uint64]
        o.orc_destroy.argtypes = [vp]
        o.orc_create_primes.argtypes = [ctypes.c_uint64, i32p, ctypes.c_int, u64p]
        o.orc_mt19937_64_fill.argtypes = [ctypes.c_uint64, ctypes.c_uint64, u64p, ctypes.c_size_t, ctypes.c_int]
        o.orc_ntt_forward.argtypes = [vp, u64p, ctypes.c_int, i32p]
        o.orc_ntt_inverse.argtypes = [vp, u64p, ctypes.c_int, i32p]
        o.orc_fnwt_1d.argtypes = [u64p, u64p, u64p, u64p, ctypes.c_uint64, ctypes.c_int, ctypes.c_int]
        o.orc_inwt_1d.argtypes = [u64p, u64p, u64p, u64p, u64p, u64p, ctypes.c_uint64, ctypes.c_int, ctypes.c_int]
        o.orc_tensor_2x2.argtypes = [vp, u64p, u64p, u64p, ctypes.c_int]
        o.orc_tensor_square_2x2.argtypes = [vp, u64p, u64p, ctypes.c_int]
        o.orc_tensor_mxn.argtypes = [vp, u64p, ctypes.c_int, u64p, ctypes.c_int, u64p, ctypes.c_int]
        for f in ("orc_poly_add", "orc_poly_sub", "orc_poly_mul"):
            getattr(o, f).argtypes = [vp, u64p, u64p, u64p, ctypes.c_int]
        o.orc_poly_negate.argtypes = [vp, u64p, u64p, ctypes.c_int]
        o.orc_beta.argtypes = [vp, ctypes.c_int]
        o.orc_modup.argtypes = [vp, ctypes.c_int, u64p, u64p]
        o.orc_inner_prod.argtypes = [vp, ctypes.c_int, u64p, u64p, u64p]
        o.orc_moddown_from_ntt.argtypes = [vp, ctypes.c_int, u64p, u64p]
        o.orc_keyswitch.argtypes = [vp, ctypes.c_int, u64p, u64p, u64p]
        o.orc_multiply_relin.argtypes = [vp, ctypes.c_int, u64p, u64p, u64p, u64p]
        o.orc_hps_aux.argtypes = [vp, u64p, i32p]
        o.orc_bfv_multiply_hps.argtypes = [vp, u64p, u64p, u64p]
        o.orc_ckks_encode.argtypes = [vp, ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.c_uint64, ctypes.c_double, u64p]
        o.orc_ckks_decode.argtypes = [vp, ctypes.c_int, u64p, ctypes.c_double, ctypes.POINTER(ctypes.c_double)]
        o.orc_prng_block.argtypes = [vp, ctypes.c_char_p, ctypes.c_uint64]
        o.orc_prng_block.restype = None
        o.orc_sample_poly.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_char_p, u64p]
        o.orc_gen_secretkey.argtypes = [vp, ctypes.c_char_p, u64p]
        o.orc_gen_secretkey.restype = None
        o.orc_encrypt_zero_symmetric.argtypes = [vp, ctypes.c_int, u64p, ctypes.c_char_p, ctypes.c_char_p, u64p]
        o.orc_encrypt_zero_asymmetric.argtypes = [vp, u64p, ctypes.c_char_p, ctypes.c_char_p, u64p]
        o.orc_gen_kswitch_key.argtypes = [vp, u64p, u64p, ctypes.c_char_p, u64p]
        o.orc_encrypt_add_plain.argtypes = [vp, ctypes.c_int, u64p, u64p]
        o.orc_plain_add.argtypes = [vp, ctypes.c_int, u64p, u64p, ctypes.c_int, ctypes.c_uint64]
        o.orc_plain_multiply.argtypes = [vp, ctypes.c_int, u64p, ctypes.c_int, u64p]
        o.orc_batch_encode.argtypes = [ctypes.c_uint64, ctypes.c_uint64, u64p, ctypes.c_uint64, u64p]
        o.orc_batch_decode.argtypes = [ctypes.c_uint64, ctypes.c_uint64, u64p, u64p]
        o.orc_decrypt.argtypes = [vp, ctypes.c_int, u64p, ctypes.c_int, u64p, ctypes.c_int, ctypes.c_uint64, u64p]
        o.orc_bfv_multiply_hps_overq.argtypes = [vp, u64p, u64p, u64p, ctypes.c_int]
        o.orc_bfv_keyswitch_leveled.argtypes = [vp, u64p, u64p, u64p, ctypes.c_int, ctypes.c_int]
        o.orc_bfv_multiply_relin_hps_overq.argtypes = [vp, u64p, u64p, u64p, u64p, ctypes.c_int]
        o.orc_bfv_multiply_relin_hps.argtypes = [vp, u64p, u64p, u64p, u64p]
        o.orc_behz_aux.argtypes = [vp, u64p, i32p]
        o.orc_bfv_multiply_behz.argtypes = [vp, u64p, u64p, u64p]
        o.orc_bfv_multiply_relin_behz.argtypes = [vp, u64p, u64p, u64p, u64p]
        o.orc_galois_table.argtypes = [ctypes.c_uint64, ctypes.c_uint32, u32p]
        o.orc_galois_elt_from_step.restype = ctypes.c_uint32
        o.orc_galois_elt_from_step.argtypes = [ctypes.c_int, ctypes.c_uint64]
        o.orc_apply_galois.argtypes = [vp, ctypes.c_int, u64p, ctypes.c_uint32, u64p]
        o.orc_apply_galois_ntt.argtypes = [vp, u64p, u64p, ctypes.c_int, u32p]
        o.orc_apply_galois_ntt.restype = None
        o.orc_apply_galois_coeff.argtypes = [vp, u64p, u64p, ctypes.c_int, ctypes.c_uint32]
        o.orc_apply_galois_coeff.restype = None
        o.orc_hoisting.argtypes = [vp, ctypes.c_int, u64p, u32p, ctypes.c_int, ctypes.POINTER(u64p)]
        o.orc_rescale.argtypes = [vp, ctypes.c_int, u64p, ctypes.c_int, u64p]
        o.orc_divide_round_q_last.argtypes = [vp, ctypes.c_int, u64p, ctypes.c_int, u64p]
        o.orc_bgv_mod_switch.argtypes = [vp, ctypes.c_int, u64p, ctypes.c_int, u64p]
        o.orc_mod_switch_drop.argtypes = [vp, ctypes.c_int, u64p, ctypes.c_int, u64p]
        for f in ("orc_twiddle", "orc_twiddle_shoup", "orc_itwiddle", "orc_itwiddle_shoup"):
            getattr(o, f).restype = u64p
            getattr(o, f).argtypes = [vp, ctypes.c_int]
        o.orc_n_inv.restype = ctypes.c_uint64
        o.orc_n_inv.argtypes = [vp, ctypes.c_int]
        o.orc_minimal_primitive_root.restype = ctypes.c_uint64
        o.orc_minimal_primitive_root.argtypes = [ctypes.c_uint64, ctypes.c_uint64]
        o.orc_shoup.restype = ctypes.c_uint64
        o.orc_shoup.argtypes = [ctypes.c_uint64, ctypes.c_uint64]
        o.orc_barrett_ratio.argtypes = [ctypes.c_uint64, u64p]
        o.orc_set_threads(min(os.cpu_count() or 1, 32))
        _oracle = o
    return _oracle


_ref = None


def reference():
    """The unmodified reference + shim (oracle/ref_shim.cu), or None when it was not built."""
    global _ref
    if _ref is None:
        if not os.path.exists(REF_SO):
            return None
        r = ctypes.CDLL(REF_SO)
        r.ref_last_error.restype = ctypes.c_char_p
        r.ref_host_create_primes.argtypes = [ctypes.c_size_t, i32p, ctypes.c_int, u64p]
        r.ref_host_ntt_table.argtypes = [ctypes.c_int, ctypes.c_uint64, u64p, u64p, u64p, u64p, u64p]
        r.ref_host_bconv_tables.argtypes = [u64p, ctypes.c_int, u64p, ctypes.c_int, u64p, u64p]
        r.ref_host_galois_elt.restype = ctypes.c_uint32
        r.ref_host_galois_elt.argtypes = [ctypes.c_int, ctypes.c_size_t]
        r.ref_create.restype = vp
        r.ref_create.argtypes = [ctypes.c_int, ctypes.c_size_t, u64p, ctypes.c_int, ctypes.c_int, ctypes.c_uint64,
                                 ctypes.c_int, i32p, ctypes.c_int, ctypes.c_double, ctypes.c_int]
        r.ref_destroy.argtypes = [vp]
        r.ref_dnum.argtypes = [vp]
        r.ref_galois_count.argtypes = [vp]
        r.ref_galois_elt_at.restype = ctypes.c_uint32
        r.ref_galois_elt_at.argtypes = [vp, ctypes.c_int]
        r.ref_key_get.argtypes = [vp, ctypes.c_int, ctypes.c_int, u64p]
        r.ref_key_set.argtypes = [vp, ctypes.c_int, ctypes.c_int, u64p]
        r.ref_ntt.argtypes = [vp, u64p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int]
        r.ref_multiply_relin.argtypes = [vp, ctypes.c_size_t, u64p, u64p, u64p]
        r.ref_multiply.argtypes = [vp, ctypes.c_size_t, u64p, u64p, u64p]
        if hasattr(r, "ref_ckks_encode"):
            r.ref_ckks_encode.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.c_size_t, ctypes.c_size_t,
                                          ctypes.c_double, u64p]
        if hasattr(r, "ref_ckks_decode"):
            r.ref_ckks_decode.argtypes = [vp, u64p, ctypes.c_size_t, ctypes.c_double, ctypes.POINTER(ctypes.c_double)]
        if hasattr(r, "ref_ckks_roundtrip"):
            r.ref_ckks_roundtrip.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.c_size_t, ctypes.c_size_t,
                                             ctypes.c_double, ctypes.POINTER(ctypes.c_double)]
        if hasattr(r, "ref_sample_poly"):
            r.ref_sample_poly.argtypes = [vp, ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, u64p]
            r.ref_encrypt.argtypes = [vp, ctypes.c_int, ctypes.c_size_t, u64p, u64p]
        if hasattr(r, "ref_load_symmetric"):
            r.ref_encrypt_save_symmetric.argtypes = [vp, ctypes.c_size_t, u64p, ctypes.c_char_p, ctypes.c_size_t, u64p]
            r.ref_encrypt_save_symmetric.restype = ctypes.c_long
            r.ref_load_symmetric.argtypes = [vp, ctypes.c_char_p, ctypes.c_size_t, u64p]
        if hasattr(r, "ref_plain_op"):
            r.ref_plain_op.argtypes = [vp, ctypes.c_int, ctypes.c_size_t, u64p, ctypes.c_size_t, u64p, ctypes.c_uint64, u64p]
            r.ref_add_sub.argtypes = [vp, ctypes.c_int, ctypes.c_size_t, u64p, u64p, ctypes.c_size_t, ctypes.c_uint64,
                                      ctypes.c_uint64, u64p, u64p]
        if hasattr(r, "ref_public_key_stream"):
            r.ref_public_key_stream.argtypes = [vp, ctypes.c_char_p, ctypes.c_size_t]
            r.ref_public_key_stream.restype = ctypes.c_long
        if hasattr(r, "ref_batch_encode"):
            r.ref_batch_encode.argtypes = [vp, u64p, ctypes.c_size_t, u64p]
            r.ref_batch_decode.argtypes = [vp, u64p, u64p]
        if hasattr(r, "ref_save_ct"):
            u8p = ctypes.POINTER(ctypes.c_ubyte)
            r.ref_save_ct.restype = ctypes.c_long
            r.ref_save_ct.argtypes = [vp, ctypes.c_size_t, u64p, ctypes.c_size_t, ctypes.c_double, ctypes.c_size_t, u8p,
                                      ctypes.c_size_t]
            r.ref_load_ct.argtypes = [u8p, ctypes.c_size_t, u64p, ctypes.POINTER(ctypes.c_size_t),
                                      ctypes.POINTER(ctypes.c_double)]
            r.ref_save_key.restype = ctypes.c_long
            r.ref_save_key.argtypes = [vp, ctypes.c_int, u8p, ctypes.c_size_t]
        if hasattr(r, "ref_decrypt"):
            r.ref_secret_key.argtypes = [vp, u64p]
            r.ref_decrypt.argtypes = [vp, ctypes.c_size_t, u64p, ctypes.c_size_t, ctypes.c_uint64, u64p]
        if hasattr(r, "ref_multiply_deg"):
            r.ref_multiply_deg.argtypes = [vp, ctypes.c_size_t, u64p, u64p, ctypes.c_size_t, ctypes.c_size_t, ctypes.c_int,
                                           u64p]
        if hasattr(r, "ref_nwt_1d"):
            r.ref_nwt_1d.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, u64p, ctypes.c_int]
        if hasattr(r, "ref_multiply_sizes"):
            r.ref_multiply_sizes.argtypes = [vp, ctypes.c_size_t, u64p, ctypes.c_size_t, u64p, ctypes.c_size_t, u64p]
        r.ref_modup.argtypes = [vp, ctypes.c_size_t, u64p, u64p]
        r.ref_inner_prod.argtypes = [vp, ctypes.c_size_t, ctypes.c_int, u64p, u64p]
        r.ref_moddown.argtypes = [vp, ctypes.c_size_t, u64p, u64p]
        r.ref_rotate.argtypes = [vp, ctypes.c_size_t, u64p, ctypes.c_int, u64p]
        r.ref_rescale.argtypes = [vp, ctypes.c_size_t, u64p, ctypes.c_size_t, u64p]
        r.ref_mod_switch.argtypes = [vp, ctypes.c_size_t, u64p, ctypes.c_size_t, u64p]
        r.ref_time_op.argtypes = [vp, ctypes.c_int, ctypes.c_size_t, u64p, u64p, ctypes.c_int, ctypes.c_int,
                                  ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
        _ref = r
    return _ref


# ---------------------------------------------------------------------------------------------------------
# parameter sets (SURVEY.md section 8 / BASELINE.md section 3)
# ---------------------------------------------------------------------------------------------------------
class ParamSet:
    def __init__(self, name, n, bit_sizes, size_P, scheme=3, t=0):
        self.name, self.n, self.bit_sizes, self.size_P, self.scheme, self.t = name, n, list(bit_sizes), size_P, scheme, t
        o = oracle()
        bits = (ctypes.c_int * len(bit_sizes))(*bit_sizes)
        primes = np.zeros(len(bit_sizes), dtype=np.uint64)
        assert o.orc_create_primes(n, bits, len(bit_sizes), P(primes)) == 0
        self.primes = primes
        self.size_QP = len(bit_sizes)
        self.size_Q = self.size_QP - size_P
        self._octx = None

    def octx(self):
        if self._octx is None:
            self._octx = oracle().orc_create(self.scheme, self.n, P(self.primes), self.size_QP, self.size_P, self.t)
            assert self._octx
        return self._octx

    def limbs(self, chain_index=1):
        return self.size_Q - (chain_index - 1)

    def beta(self, chain_index=1):
        return -(-self.limbs(chain_index) // self.size_P)

    def row(self, l, j):
        return j if j < l else self.size_Q + (j - l)


def params_c1():  # config 1: fwd+inv NTT, N=2^12, one 50-bit prime (test/ntt_test.cu:78 pattern)
    return ParamSet("C1", 4096, [50], 0)


def params_primary():  # N=2^16, L=16: {60, 40x15, 60x4}, alpha=4
    return ParamSet("primary", 65536, [60] + [40] * 15 + [60] * 4, 4)


def params_secondary():  # {60, 40x15, 60}, alpha=1 (single-P fast paths)
    return ParamSet("secondary", 65536, [60] + [40] * 15 + [60], 1)


def params_bfv_bench(which=0):
    """BFV parameter sets of benchmark/bfv_bench.cu:303-345 (N=2^14, log QP = 438), t = PlainModulus::Batching(n, 20)"""
    n = 16384
    sets = [([54] * 7 + [60], 1), ([36] * 11 + [42], 1), ([36] * 8 + [37, 37, 38, 38], 4)]
    bits, size_P = sets[which]
    o = oracle()
    t = np.zeros(1, dtype=np.uint64)
    assert o.orc_create_primes(n, (ctypes.c_int * 1)(20), 1, P(t)) == 0
    return ParamSet(f"bfv14_{which}", n, bits, size_P, scheme=2, t=int(t[0]))


def params_small(n=4096, l=5, alpha=2, qbits=40, pbits=50, scheme=3, t=0):  # oracle-in-seconds sizes
    return ParamSet(f"small{n}", n, [qbits + 10] + [qbits] * (l - 1) + [pbits] * alpha, alpha, scheme, t)


# ---------------------------------------------------------------------------------------------------------
# synthetic inputs: every word uniform in [0, q_limb) from std::mt19937_64(seed) with rejection
# ---------------------------------------------------------------------------------------------------------
def uniform_limbs(ps, rows, seed, polys=1):
    """[polys][len(rows)][n] residues; limb i is reduced modulo primes[rows[i]]."""
    o = oracle()
    out = np.zeros((polys, len(rows), ps.n), dtype=np.uint64)
    for p in range(polys):
        for i, r in enumerate(rows):
            o.orc_mt19937_64_fill(seed * 1000003 + p * 1009 + i, int(ps.primes[r]), P(out[p, i]), ps.n, 1)
    return out


def ciphertext(ps, seed, chain_index=1, polys=2):
    l = ps.limbs(chain_index)
    return uniform_limbs(ps, list(range(l)), seed, polys)


def switch_key(ps, seed):
    """[dnum][2][size_QP][n] (PhantomRelinKey layout, include/secretkey.h:102-127)"""
    dnum = ps.beta(1)
    out = np.zeros((dnum, 2, ps.size_QP, ps.n), dtype=np.uint64)
    for d in range(dnum):
        out[d] = uniform_limbs(ps, list(range(ps.size_QP)), seed + d, 2)
    return out


def edge_vectors(ps, rows):
    """all-0, all-(q-1), unit impulses at {0, 1, N/2, N-1}"""
    n = ps.n
    vecs = []
    z = np.zeros((len(rows), n), dtype=np.uint64)
    vecs.append(z.copy())
    m = z.copy()
    for i, r in enumerate(rows):
        m[i, :] = ps.primes[r] - np.uint64(1)
    vecs.append(m)
    for pos in (0, 1, n // 2, n - 1):
        d = z.copy()
        d[:, pos] = 1
        vecs.append(d)
    return vecs

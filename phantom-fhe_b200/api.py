"""Host-side mirror of the reference's interface for the hot path.

Names, argument meaning and error behaviour follow the reference (paths relative to the reference root):
EncryptionParameters (include/host/encryptionparams.h), CoeffModulus::Create (src/host/modulus.cu:79-110),
PhantomContext (include/context.cuh), PhantomCiphertext (include/ciphertext.h), PhantomRelinKey /
PhantomGaloisKey (include/secretkey.h:99-219) and the evaluator free functions (include/evaluate.cuh:37-245).
Where the reference throws std::invalid_argument this raises ValueError with the same message;
std::logic_error -> RuntimeError.  All arithmetic happens in libpfhe_b200.so on the current CUDA stream.
"""
import copy
import ctypes
import enum
import math
import os

import numpy as np
import torch

from . import serial
from ._lib import lib, check, u64p, u32p, i32p


class scheme_type(enum.IntEnum):  # host/encryptionparams.h:14-22
    none = 0
    bgv = 1
    bfv = 2
    ckks = 3


class mul_tech_type(enum.IntEnum):  # host/encryptionparams.h:25-35
    none = 0
    behz = 1
    hps = 2
    hps_overq = 3
    hps_overq_leveled = 4


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr())


class CoeffModulus:
    @staticmethod
    def Create(poly_modulus_degree, bit_sizes):
        """CoeffModulus::Create (src/host/modulus.cu:79-110)."""
        bits = (ctypes.c_int * len(bit_sizes))(*bit_sizes)
        out = (ctypes.c_uint64 * len(bit_sizes))()
        check(lib.pfhe_create_primes(poly_modulus_degree, bits, len(bit_sizes), out))
        return [int(v) for v in out]


class PlainModulus:
    @staticmethod
    def Batching(poly_modulus_degree, bit_size):
        """PlainModulus::Batching (include/host/modulus.h:317-319): the first prime of that size that is 1 mod 2N"""
        return CoeffModulus.Create(poly_modulus_degree, [bit_size])[0]


# names of the reference's Python binding (python/src/binding.cu:41-43)
create_coeff_modulus = CoeffModulus.Create
create_plain_modulus = PlainModulus.Batching


def get_elt_from_step(step, coeff_count):
    """include/galois.cuh:16-49"""
    elt = ctypes.c_uint32()
    check(lib.pfhe_galois_elt_from_step(int(step), int(coeff_count), ctypes.byref(elt)))
    return int(elt.value)


def get_elts_from_steps(steps, coeff_count):
    return [get_elt_from_step(s, coeff_count) for s in steps]


class EncryptionParameters:
    """include/host/encryptionparams.h:57-150 (the setters the hot path depends on)."""

    def __init__(self, scheme=scheme_type.none):
        self.scheme = scheme_type(scheme)
        self.poly_modulus_degree = 0
        self.coeff_modulus = []
        self.special_modulus_size = 1  # reference default, encryptionparams.h:235
        self.galois_elts = []
        self.plain_modulus = 0
        # reference default for BFV is HPS (encryptionparams.h:41-47)
        self.mul_tech = mul_tech_type.hps if self.scheme == scheme_type.bfv else mul_tech_type.none

    def set_mul_tech(self, mul_tech):
        if self.scheme != scheme_type.bfv:
            raise ValueError("mul_tech selection is only supported for BFV")
        if mul_tech_type(mul_tech) == mul_tech_type.none:
            raise ValueError("unsupported multiplication technique for BFV")
        self.mul_tech = mul_tech_type(mul_tech)

    def set_poly_modulus_degree(self, n):
        if self.scheme == scheme_type.none and n:
            raise RuntimeError("poly_modulus_degree is not supported for this scheme")
        self.poly_modulus_degree = int(n)

    def set_coeff_modulus(self, primes):
        if self.scheme == scheme_type.none and primes:
            raise RuntimeError("coeff_modulus is not supported for this scheme")
        self.coeff_modulus = [int(p) for p in primes]

    def set_special_modulus_size(self, k):
        self.special_modulus_size = int(k)

    def set_galois_elts(self, elts):
        self.galois_elts = [int(e) for e in elts]

    def set_plain_modulus(self, t):
        if self.scheme not in (scheme_type.bfv, scheme_type.bgv) and t:
            raise RuntimeError("plain_modulus is not supported for this scheme")
        self.plain_modulus = int(t)


class PhantomContext:
    """Constant tables of the hot path for one parameter set (reference src/context.cu:121-232)."""

    def __init__(self, params):
        if not torch.cuda.is_available():
            raise RuntimeError("CUDA Runtime Error: no CUDA device (phantom-fhe_b200 has no CPU path)")
        self.parms = params
        self.poly_degree = params.poly_modulus_degree
        self.size_QP = len(params.coeff_modulus)
        self.size_P = params.special_modulus_size
        self.size_Q = self.size_QP - self.size_P
        self.scheme = params.scheme
        primes = (ctypes.c_uint64 * self.size_QP)(*params.coeff_modulus)
        elts = (ctypes.c_uint32 * max(1, len(params.galois_elts)))(*params.galois_elts)
        handle = ctypes.c_void_p()
        check(lib.pfhe_engine_create(ctypes.byref(handle), int(params.scheme), self.poly_degree, primes, self.size_QP,
                                     self.size_P, params.plain_modulus, elts, len(params.galois_elts)))
        self._h = handle
        # the key order of the context: the given elements, or the reference's default set when none were given
        # (PhantomGaloisTool, include/galois.cuh:84-89); the mirror's key generation and lookups index this list
        cnt = lib.pfhe_galois_elts(handle, None, 0)
        buf = (ctypes.c_uint32 * max(1, cnt))()
        lib.pfhe_galois_elts(handle, buf, cnt)
        self.parms = copy.copy(params)
        self.parms.galois_elts = [int(buf[i]) for i in range(cnt)]
        if self.scheme == scheme_type.bfv:
            check(lib.pfhe_engine_set_mul_tech(handle, int(params.mul_tech)))
        self.device = torch.device("cuda", torch.cuda.current_device())

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.pfhe_engine_destroy(h)
            self._h = None

    def get_first_index(self):
        return 1

    def coeff_modulus_size(self, chain_index):
        if chain_index < 1 or chain_index > self.size_Q:
            raise ValueError("index is invalid!")
        return self.size_Q - (chain_index - 1)

    def dnum(self, chain_index=1):
        return lib.pfhe_dnum(self._h, chain_index)

    def launch_count(self):
        return int(lib.pfhe_launch_count(self._h))


def _to_dev(host_u64, device):
    a = np.ascontiguousarray(host_u64, dtype=np.uint64)
    return torch.from_numpy(a.view(np.int64)).to(device)


class PhantomCiphertext:
    """[size][coeff_modulus_size][poly_modulus_degree] uint64 words on the device (include/ciphertext.h:15-25)."""

    def __init__(self, context=None, data=None, chain_index=1, scale=1.0, is_ntt_form=True):
        self.data = data
        self.chain_index = chain_index
        self.scale = scale
        self.is_ntt_form = is_ntt_form
        self.correction_factor = 1
        self.noise_scale_deg = 1     # noiseScaleDeg_ (include/ciphertext.h:21): read by mul_tech hps_overq_leveled
        self.is_asymmetric = False   # is_asymmetric_ (:23)
        self.seed = None             # seed_ of c1 (symmetric encryption; save_symmetric writes it instead of c1)

    @classmethod
    def from_host(cls, context, words, chain_index=1, scale=1.0, is_ntt_form=True):
        l = context.coeff_modulus_size(chain_index)
        words = np.asarray(words, dtype=np.uint64).reshape(-1, l, context.poly_degree)
        return cls(context, _to_dev(words, context.device), chain_index, scale, is_ntt_form)

    def to_host(self):
        torch.cuda.current_stream().synchronize()
        return self.data.cpu().numpy().view(np.uint64)

    def size(self):
        return 0 if self.data is None else self.data.shape[0]

    def save(self, stream):
        """PhantomCiphertext::save (include/ciphertext.h:173-190): byte-identical stream."""
        serial.write_ciphertext(stream, self.to_host(), self.chain_index, self.scale, self.correction_factor,
                                self.noise_scale_deg, self.is_ntt_form, self.is_asymmetric)

    @classmethod
    def load(cls, context, stream):
        """PhantomCiphertext::load (include/ciphertext.h:192-213)."""
        words, h = serial.read_ciphertext(stream)
        if words.shape[2] != context.poly_degree or words.shape[1] != context.coeff_modulus_size(h["chain_index"]):
            raise ValueError("ciphertext stream does not belong to this context")
        c = cls(context, _to_dev(words, context.device), h["chain_index"], h["scale"], h["is_ntt_form"])
        c.correction_factor, c.noise_scale_deg, c.is_asymmetric = h["correction_factor"], h["noise_scale_deg"], h["is_asymmetric"]
        return c

    def save_symmetric(self, stream):
        """PhantomCiphertext::save_symmetric (include/ciphertext.h:216-245): c0 and the seed of c1, half the bytes."""
        if self.is_asymmetric or getattr(self, "seed", None) is None:
            raise RuntimeError("Asymmetric ciphertext does not have seed.")
        if self.size() != 2:
            raise RuntimeError("This method is only for 2-polynomial ciphertext.")
        serial.write_ciphertext_symmetric(stream, self.to_host()[0], self.seed, self.chain_index, self.scale,
                                          self.correction_factor, self.noise_scale_deg, self.is_ntt_form)

    @classmethod
    def load_symmetric(cls, context, stream):
        """PhantomCiphertext::load_symmetric (include/ciphertext.h:247-307): c1 is drawn again from the seed on the device
        (sample_uniform_poly), and brought to coefficient form for BFV.  First data level only, like the reference."""
        c0, seed, h = serial.read_ciphertext_symmetric(stream)
        l, n = c0.shape
        if (n != context.poly_degree or l != context.coeff_modulus_size(context.get_first_index())
                or h["chain_index"] != context.get_first_index()):
            raise RuntimeError("Only support ciphertext without modulus switching.")
        data = torch.empty((2, l, n), dtype=torch.int64, device=context.device)
        data[0].copy_(_to_dev(c0, context.device))
        check(lib.pfhe_sample_poly(context._h, 2, l, seed, _ptr(data[1]), _stream()))
        if not h["is_ntt_form"]:
            check(lib.pfhe_ntt_backward_inplace(context._h, _ptr(data[1]), l, 0, _stream()))
        c = cls(context, data, h["chain_index"], h["scale"], h["is_ntt_form"])
        c.correction_factor, c.noise_scale_deg, c.seed = h["correction_factor"], h["noise_scale_deg"], seed
        return c

    def coeff_modulus_size(self):
        return self.data.shape[1]

    def clone(self):
        c = PhantomCiphertext(None, self.data.clone(), self.chain_index, self.scale, self.is_ntt_form)
        c.correction_factor, c.noise_scale_deg, c.is_asymmetric = self.correction_factor, self.noise_scale_deg, self.is_asymmetric
        return c


class PhantomCKKSEncoder:
    """PhantomCKKSEncoder (include/ckks.h, src/ckks.cu:66-190): encode and decode."""

    def __init__(self, context):
        if context.scheme != scheme_type.ckks:
            raise ValueError("unsupported scheme")
        self._slots = context.poly_degree >> 1

    def slot_count(self):
        return self._slots

    def encode(self, context, values, scale, chain_index=1):
        """complex (or real) slot values -> device plaintext [l][N] in NTT form at chain_index"""
        v = np.ascontiguousarray(np.asarray(values, dtype=np.complex128))
        if v.size == 0:
            raise ValueError("Input vector is empty")
        if v.size > self._slots:
            raise ValueError("Input vector exceeds max slots")
        d_in = torch.from_numpy(v.view(np.float64).copy()).to(context.device)
        l = context.coeff_modulus_size(chain_index)
        plain = torch.empty((l, context.poly_degree), dtype=torch.int64, device=context.device)
        check(lib.pfhe_ckks_encode(context._h, chain_index, _ptr(d_in), v.size, float(scale), _ptr(plain), _stream()))
        return plain

    def decode(self, context, plain, scale, chain_index=None):
        """device plaintext [l][N] in NTT form with the given scale -> numpy complex128 [N/2] (the reference reads
        chain_index and scale off the PhantomPlaintext; here the level defaults to the one with l limbs)"""
        if plain.dim() != 2 or plain.shape[1] != context.poly_degree:
            raise ValueError("plaintext shape")
        if chain_index is None:
            chain_index = context.size_Q - plain.shape[0] + 1
        if context.coeff_modulus_size(chain_index) != plain.shape[0]:
            raise ValueError("plaintext does not match chain_index")
        out = torch.empty(self._slots * 2, dtype=torch.float64, device=plain.device)
        check(lib.pfhe_ckks_decode(context._h, chain_index, _ptr(plain.contiguous()), float(scale), _ptr(out), _stream()))
        torch.cuda.current_stream().synchronize()
        return out.cpu().numpy().view(np.complex128)

    # names of the reference's Python binding (python/src/binding.cu:98-117)
    def encode_complex_vector(self, context, values, scale, chain_index=1):
        return self.encode(context, values, scale, chain_index)

    def encode_double_vector(self, context, values, scale, chain_index=1):
        return self.encode(context, np.asarray(values, dtype=np.float64), scale, chain_index)

    def decode_complex_vector(self, context, plain, scale, chain_index=None):
        return self.decode(context, plain, scale, chain_index)

    def decode_double_vector(self, context, plain, scale, chain_index=None):
        """the real parts (decode_internal's std::vector<double> form, include/ckks.h:46-55)"""
        return np.ascontiguousarray(self.decode(context, plain, scale, chain_index).real)


def save_plaintext(stream, plain, chain_index=0, scale=1.0):
    """PhantomPlaintext::save (include/plaintext.h:69-81) of a device plaintext: CKKS [l][N] with its level and scale,
    BFV / BGV [N] (chain_index 0, scale 1 as the batch encoder leaves them)."""
    torch.cuda.current_stream().synchronize()
    serial.write_plaintext(stream, plain.cpu().numpy().view(np.uint64), chain_index, scale)


def load_plaintext(context, stream):
    """PhantomPlaintext::load (include/plaintext.h:83-97) -> (device words, chain_index, scale).  The reference's CKKS
    encoder sizes every plaintext for the full chain and fills the first l limbs (include/ckks.h:78-80), so a stream written
    by a stock build may carry more limbs than the level has: the limbs of the level are kept."""
    words, chain_index, scale = serial.read_plaintext(stream)
    if words.shape[1] != context.poly_degree:
        raise ValueError("plaintext stream does not belong to this context")
    if context.scheme == scheme_type.ckks:
        l = context.coeff_modulus_size(chain_index)
        if words.shape[0] < l:
            raise ValueError("plaintext stream does not belong to this context")
        return _to_dev(words[:l], context.device), chain_index, scale
    return _to_dev(words[0], context.device), chain_index, scale


class PhantomBatchEncoder:
    """PhantomBatchEncoder (include/batchencoder.h, src/batchencoder.cu): BFV / BGV slot packing over the plain modulus."""

    def __init__(self, context):
        if context.scheme not in (scheme_type.bfv, scheme_type.bgv):
            raise ValueError("PhantomBatchEncoder only supports BFV/BGV scheme")
        self._slots = context.poly_degree

    def slot_count(self):
        return self._slots

    def encode(self, context, values_matrix):
        """-> device plaintext [N] (coefficient form, residues mod t)"""
        v = np.asarray(values_matrix)
        if v.size > self._slots:
            raise RuntimeError("values_matrix size is too large")
        if v.dtype.kind == "i":
            v = v.astype(np.int64).view(np.uint64)   # negative values go down as two's complement, like the reference's
        v = np.ascontiguousarray(v, dtype=np.uint64)
        d_in = _to_dev(v, context.device) if v.size else torch.zeros(1, dtype=torch.int64, device=context.device)
        plain = torch.empty(self._slots, dtype=torch.int64, device=context.device)
        check(lib.pfhe_batch_encode(context._h, _ptr(d_in), v.size, _ptr(plain), _stream()))
        return plain

    def decode(self, context, plain):
        """-> numpy uint64 [N] slot values"""
        out = torch.empty(self._slots, dtype=torch.int64, device=plain.device)
        check(lib.pfhe_batch_decode(context._h, _ptr(plain), _ptr(out), _stream()))
        torch.cuda.current_stream().synchronize()
        return out.cpu().numpy().view(np.uint64)


def random_bytes(count=64):
    """random_bytes (include/prng.cuh:10-32): the seeds of the device generator.  The reference draws them from
    std::random_device; os.urandom is the same kind of source."""
    return os.urandom(count)


def _seed(seed):
    seed = random_bytes() if seed is None else bytes(seed)
    if len(seed) != 64:
        raise ValueError("a seed is 64 bytes (prng_seed_byte_count)")
    return seed


def _first_plain_level(context, plain):
    if context.scheme == scheme_type.ckks:
        if plain.dim() != 2 or plain.shape[1] != context.poly_degree:
            raise ValueError("CKKS plaintext is [l][N] in NTT form")
        return context.size_Q - plain.shape[0] + 1
    if plain.dim() != 1 or plain.shape[0] != context.poly_degree:
        raise ValueError("BFV / BGV plaintext is [N] residues mod t")
    return context.get_first_index()


class PhantomPublicKey:
    """PhantomPublicKey (include/secretkey.h:25-100): an encryption of zero at the key level, [2][size_QP][N] in NTT form,
    and encrypt_asymmetric (src/secretkey.cu:130-190)."""

    def __init__(self, context, pk):
        self.pk = pk   # device [2][size_QP][N]

    def save(self, stream):
        """PhantomPublicKey::save (include/secretkey.h:85-90)"""
        torch.cuda.current_stream().synchronize()
        serial.write_public_key(stream, self.pk.cpu().numpy().view(np.uint64))

    @classmethod
    def load(cls, context, stream):
        words = serial.read_public_key(stream)
        if words.shape[1] != context.size_QP or words.shape[2] != context.poly_degree:
            raise ValueError("public key stream does not belong to this context")
        return cls(context, _to_dev(words, context.device))

    def encrypt_asymmetric(self, context, plain, scale=1.0, seeds=None):
        """plain: device plaintext (BFV / BGV: [N] mod t, first data level; CKKS: [l][N] NTT form, l = size_Q: the
        reference's mod-down step serves the first data level only).  seeds = (seed_u, seed_e) or None for fresh ones."""
        chain_index = _first_plain_level(context, plain)
        su, se = (None, None) if seeds is None else seeds
        l, n = context.coeff_modulus_size(chain_index), context.poly_degree
        data = torch.empty((2, l, n), dtype=torch.int64, device=context.device)
        check(lib.pfhe_encrypt_zero_asymmetric(context._h, chain_index, _ptr(self.pk), _seed(su), _seed(se), _ptr(data),
                                               _stream()))
        check(lib.pfhe_encrypt_add_plain(context._h, chain_index, _ptr(data), _ptr(plain.contiguous()), _stream()))
        ct = PhantomCiphertext(context, data, chain_index, scale if context.scheme == scheme_type.ckks else 1.0,
                               context.scheme != scheme_type.bfv)
        ct.is_asymmetric = True
        return ct


class PhantomSecretKey:
    """PhantomSecretKey (include/secretkey.h:226-338, src/secretkey.cu:196-723): the secret key's powers in NTT form at the
    key level (secret_key_array()), key generation, symmetric encryption and decrypt().  PhantomSecretKey(context) draws a
    new key like the reference's constructor; PhantomSecretKey(context, s) wraps an existing first power (from
    PhantomSecretKey::save of a stock build, or from the caller)."""

    def __init__(self, context, secret_key_ntt=None, seed=None):
        if secret_key_ntt is None:
            self._pow = torch.empty((1, context.size_QP, context.poly_degree), dtype=torch.int64, device=context.device)
            check(lib.pfhe_gen_secretkey(context._h, _seed(seed), _ptr(self._pow), _stream()))   # gen_secretkey :345-378
            return
        s1 = np.asarray(secret_key_ntt, dtype=np.uint64).reshape(1, context.size_QP, context.poly_degree)
        self._pow = _to_dev(s1, context.device)   # [sk_max_power][size_QP][n]

    def _encrypt_zero_symmetric(self, context, chain_index, seed_a, seed_e):
        limbs = context.size_QP if chain_index == 0 else context.coeff_modulus_size(chain_index)
        ct = torch.empty((2, limbs, context.poly_degree), dtype=torch.int64, device=context.device)
        check(lib.pfhe_encrypt_zero_symmetric(context._h, chain_index, _ptr(self._pow), _seed(seed_a), _seed(seed_e), _ptr(ct),
                                              _stream()))
        return ct

    def gen_publickey(self, context, seeds=None):
        """gen_publickey (src/secretkey.cu:380-392): encrypt_zero_symmetric at the key level.  seeds = (seed_a, seed_e)"""
        sa, se = (None, None) if seeds is None else seeds
        return PhantomPublicKey(context, self._encrypt_zero_symmetric(context, 0, sa, se))

    def _kswitch_key(self, context, new_key, seeds):
        """generate_one_kswitch_key (src/secretkey.cu:297-343) -> PhantomRelinKey"""
        if context.size_P < 1 or context.size_Q % context.size_P:
            raise ValueError("size_Q must be a multiple of size_P")
        dnum = context.size_Q // context.size_P
        if seeds is None:
            seeds = b"".join(random_bytes() for _ in range(2 * dnum))
        if len(seeds) != 128 * dnum:
            raise ValueError("a key-switching key takes dnum pairs of 64-byte seeds")
        digits = [torch.empty((2, context.size_QP, context.poly_degree), dtype=torch.int64, device=context.device)
                  for _ in range(dnum)]
        ptrs = (ctypes.c_void_p * dnum)(*[d.data_ptr() for d in digits])
        check(lib.pfhe_gen_kswitch_key(context._h, _ptr(new_key), _ptr(self._pow), bytes(seeds), ptrs, _stream()))
        return PhantomRelinKey.from_device(context, digits)

    def gen_relinkey(self, context, seeds=None):
        """gen_relinkey (src/secretkey.cu:394-418): key-switching key for s^2"""
        self._compute_secret_key_array(context, 2)
        return self._kswitch_key(context, self._pow[1], seeds)

    def create_galois_keys(self, context, seeds=None):
        """create_galois_keys (src/secretkey.cu:420-461): one key-switching key per Galois element of the context, for the
        secret key under that automorphism.  seeds: one bytes object of dnum seed pairs per element, or None"""
        keys = []
        rotated = torch.empty_like(self._pow[0])
        for i, elt in enumerate(context.parms.galois_elts):
            check(lib.pfhe_galois_secret_key(context._h, _ptr(self._pow), int(elt), _ptr(rotated), _stream()))
            keys.append(self._kswitch_key(context, rotated, None if seeds is None else seeds[i]))
        gk = PhantomGaloisKey.__new__(PhantomGaloisKey)
        gk.relin_keys = keys
        return gk

    def encrypt_symmetric(self, context, plain, scale=1.0, seeds=None):
        """encrypt_symmetric (src/secretkey.cu:463-530).  plain as in PhantomPublicKey.encrypt_asymmetric, CKKS at any level.
        seeds = (seed_a, seed_e); the ciphertext keeps seed_a like the reference's seed_ptr()"""
        chain_index = _first_plain_level(context, plain)
        sa, se = (None, None) if seeds is None else seeds
        sa = _seed(sa)
        data = self._encrypt_zero_symmetric(context, chain_index, sa, se)
        check(lib.pfhe_encrypt_add_plain(context._h, chain_index, _ptr(data), _ptr(plain.contiguous()), _stream()))
        ct = PhantomCiphertext(context, data, chain_index, scale if context.scheme == scheme_type.ckks else 1.0,
                               context.scheme != scheme_type.bfv)
        ct.seed = sa
        return ct

    def secret_key_array(self):
        return self._pow

    def save(self, stream):
        """PhantomSecretKey::save (include/secretkey.h:346-364): every power computed so far."""
        torch.cuda.current_stream().synchronize()
        serial.write_secret_key(stream, self._pow.cpu().numpy().view(np.uint64))

    @classmethod
    def load(cls, context, stream):
        p = serial.read_secret_key(stream)
        if p.shape[1] != context.size_QP or p.shape[2] != context.poly_degree:
            raise ValueError("secret key stream does not belong to this context")
        k = cls(context, p[0])
        k._pow = _to_dev(p, context.device)
        return k

    def _compute_secret_key_array(self, context, max_power):
        """compute_secret_key_array (src/secretkey.cu:196-230): s^k = s^(k-1) * s, limb-wise, all key-level limbs."""
        while self._pow.shape[0] < max_power:
            nxt = torch.empty_like(self._pow[:1])
            check(lib.pfhe_multiply_rns_poly(context._h, _ptr(self._pow[-1]), _ptr(self._pow[0]), _ptr(nxt),
                                             context.size_QP, _stream()))
            self._pow = torch.cat([self._pow, nxt])

    def decrypt(self, context, cipher):
        """PhantomSecretKey::decrypt (src/secretkey.cu:693-723): returns the plaintext words on the device -- CKKS:
        [l][n] in NTT form (what the decoder takes); BFV / BGV: [n] residues mod t."""
        _require_ntt(context, cipher)
        size = cipher.size()
        self._compute_secret_key_array(context, max(1, size - 1))
        l, n = cipher.coeff_modulus_size(), context.poly_degree
        shape = (l, n) if context.scheme == scheme_type.ckks else (n,)
        out = torch.empty(shape, dtype=torch.int64, device=cipher.data.device)
        cf = cipher.correction_factor if context.scheme == scheme_type.bgv else 1
        check(lib.pfhe_decrypt(context._h, cipher.chain_index, _ptr(cipher.data), size, _ptr(self._pow), cf, _ptr(out),
                               _stream()))
        return out


class PhantomRelinKey:
    """dnum device buffers [2][size_QP][N] in NTT form + a device array of their addresses
    (include/secretkey.h:102-127, public_keys_ptr())."""

    def __init__(self, context, digits_host):
        self.digits = [_to_dev(np.asarray(d, dtype=np.uint64).reshape(2, context.size_QP, context.poly_degree),
                               context.device) for d in digits_host]
        ptrs = np.array([d.data_ptr() for d in self.digits], dtype=np.uint64)
        self._ptrs = _to_dev(ptrs, context.device)

    @classmethod
    def from_device(cls, context, digits):
        """wrap digits already on the device (key generation)"""
        k = cls.__new__(cls)
        k.digits = list(digits)
        k._ptrs = _to_dev(np.array([d.data_ptr() for d in k.digits], dtype=np.uint64), context.device)
        return k

    def public_keys_ptr(self):
        return _ptr(self._ptrs)

    def _host_digits(self):
        torch.cuda.current_stream().synchronize()
        return [d.cpu().numpy().view(np.uint64) for d in self.digits]

    def save(self, stream):
        """PhantomRelinKey::save (include/secretkey.h:129-140)."""
        serial.write_relin_key(stream, self._host_digits())

    @classmethod
    def load(cls, context, stream):
        digits = serial.read_relin_key(stream)
        if len(digits) != context.dnum(1):   # the inner product walks dnum digit pointers
            raise ValueError("relinearisation key stream does not belong to this context")
        return cls(context, digits)


class PhantomGaloisKey:
    """Relin keys indexed like the context's Galois elements (include/secretkey.h:168-192)."""

    def __init__(self, context, keys_by_elt):
        self.relin_keys = [PhantomRelinKey(context, digits) for digits in keys_by_elt]

    def get_relin_keys(self, index):
        return self.relin_keys[index]

    def save(self, stream):
        """PhantomGaloisKey::save (include/secretkey.h:194-205)."""
        serial.write_galois_key(stream, [k._host_digits() for k in self.relin_keys])

    @classmethod
    def load(cls, context, stream):
        return cls(context, serial.read_galois_key(stream))


# ---------------------------------------------------------------------------------------------------------
# evaluator (include/evaluate.cuh:37-245)
# ---------------------------------------------------------------------------------------------------------
def _require_ntt(context, ct):
    if context.scheme in (scheme_type.ckks, scheme_type.bgv) and not ct.is_ntt_form:
        name = "CKKS" if context.scheme == scheme_type.ckks else "BGV"
        raise ValueError(f"{name} encrypted must be in NTT form")
    if context.scheme == scheme_type.bfv and ct.is_ntt_form:
        raise ValueError("BFV encrypted cannot be in NTT form")


def _leveled(context):
    return context.scheme == scheme_type.bfv and context.parms.mul_tech == mul_tech_type.hps_overq_leveled


def _levels_to_drop(context, depth, is_key_switch, is_asymmetric):
    """FindLevelsToDrop (src/evaluate.cu:550-643)."""
    lv = ctypes.c_int()
    check(lib.pfhe_find_levels_to_drop(context._h, depth, int(is_key_switch), int(is_asymmetric), ctypes.byref(lv)))
    return lv.value


def multiply_inplace(context, encrypted1, encrypted2):
    """multiply_inplace (src/evaluate.cu:1029-1057 -> bgv_ckks_multiply :345-397, bfv_multiply_behz :451-548)."""
    _require_ntt(context, encrypted1)
    _require_ntt(context, encrypted2)
    if encrypted1.chain_index != encrypted2.chain_index:   # the reference's checks, in its order (evaluate.cu:1033-1040)
        raise ValueError("encrypted1 and encrypted2 parameter mismatch")
    if encrypted1.is_ntt_form != encrypted2.is_ntt_form:
        raise ValueError("NTT form mismatch")
    if not _are_close(encrypted1.scale, encrypted2.scale):
        raise ValueError("scale mismatch")
    if encrypted1.size() != encrypted2.size():
        raise ValueError("poly number mismatch")
    l, n = encrypted1.coeff_modulus_size(), context.poly_degree
    s1, s2 = encrypted1.size(), encrypted2.size()
    dst = torch.empty((s1 + s2 - 1, l, n), dtype=torch.int64, device=encrypted1.data.device)
    a, b = encrypted1.data, encrypted2.data
    if _leveled(context) and encrypted1.chain_index == 1:   # bfv_multiply_hps, leveled branch (evaluate.cu:680-690, 798-800)
        if s1 != 2 or s2 != 2:
            raise RuntimeError("dest_size must be 3 when computing BFV multiplication using HPS")
        deg = max(encrypted1.noise_scale_deg, encrypted2.noise_scale_deg)
        drop = _levels_to_drop(context, deg - 1, False, encrypted1.is_asymmetric)
        check(lib.pfhe_multiply_leveled(context._h, _ptr(a), _ptr(b), _ptr(dst), drop, _stream()))
        encrypted1.noise_scale_deg = deg + 1
    elif s1 == 2 and s2 == 2:
        check(lib.pfhe_multiply(context._h, encrypted1.chain_index, _ptr(a), _ptr(a if encrypted1 is encrypted2 else b),
                                _ptr(dst), _stream()))
    else:   # tensor_prod_mxn_rns_poly branch of bgv_ckks_multiply (evaluate.cu:382-386)
        check(lib.pfhe_multiply_sizes(context._h, encrypted1.chain_index, _ptr(a), s1, _ptr(b), s2, _ptr(dst), _stream()))
    encrypted1.data = dst
    _after_product(context, encrypted1, encrypted2)


def _after_product(context, encrypted1, encrypted2):
    """bookkeeping of bgv_ckks_multiply (src/evaluate.cu:388-396): CKKS scales multiply, BGV correction factors multiply mod t"""
    if context.scheme == scheme_type.ckks:
        encrypted1.scale = encrypted1.scale * encrypted2.scale
    elif context.scheme == scheme_type.bgv:
        encrypted1.correction_factor = encrypted1.correction_factor * encrypted2.correction_factor % context.parms.plain_modulus


def relinearize_inplace(context, encrypted, relin_keys):
    """relinearize_inplace (src/evaluate.cu:1342-1374)."""
    if encrypted.size() != 3:
        raise ValueError("destination_size must be 3")
    _require_ntt(context, encrypted)
    if _leveled(context) and encrypted.chain_index == 1:   # keyswitch_inplace, leveled branch (eval_key_switch.cu:113-123)
        drop = _levels_to_drop(context, encrypted.noise_scale_deg - 1, False, encrypted.is_asymmetric)
        check(lib.pfhe_keyswitch_leveled_inplace(context._h, _ptr(encrypted.data), _ptr(encrypted.data[2]),
                                                 relin_keys.public_keys_ptr(), drop, _stream()))
    else:
        check(lib.pfhe_relinearize_inplace(context._h, encrypted.chain_index, _ptr(encrypted.data),
                                           relin_keys.public_keys_ptr(), _stream()))
    encrypted.data = encrypted.data[:2]


def multiply_and_relin_inplace(context, encrypted1, encrypted2, relin_keys):
    """multiply_and_relin_inplace (src/evaluate.cu:1061-1104), fused tensor + key-switch."""
    _require_ntt(context, encrypted1)
    _require_ntt(context, encrypted2)
    if encrypted1.chain_index != encrypted2.chain_index:
        raise ValueError("encrypted1 and encrypted2 parameter mismatch")
    if encrypted1.is_ntt_form != encrypted2.is_ntt_form:
        raise ValueError("NTT form mismatch")
    if not _are_close(encrypted1.scale, encrypted2.scale):
        raise ValueError("scale mismatch")
    if encrypted1.size() != encrypted2.size():
        raise ValueError("poly number mismatch")
    if encrypted1.size() != 2:   # the reference's relinearisation step refuses anything but a 3-polynomial product
        raise ValueError("destination_size must be 3")
    l, n = encrypted1.coeff_modulus_size(), context.poly_degree
    dst = torch.empty((2, l, n), dtype=torch.int64, device=encrypted1.data.device)
    if _leveled(context) and encrypted1.chain_index == 1:   # bfv_mul_relin_hps, leveled branch (evaluate.cu:845-856, 962-964)
        deg = max(encrypted1.noise_scale_deg, encrypted2.noise_scale_deg)
        drop = _levels_to_drop(context, deg - 1, False, encrypted1.is_asymmetric)
        check(lib.pfhe_multiply_and_relin_leveled(context._h, _ptr(encrypted1.data), _ptr(encrypted2.data), _ptr(dst),
                                                  relin_keys.public_keys_ptr(), drop, _stream()))
        encrypted1.noise_scale_deg = deg + 1
    else:
        check(lib.pfhe_multiply_and_relin(context._h, encrypted1.chain_index, _ptr(encrypted1.data),
                                          _ptr(encrypted2.data), _ptr(dst), relin_keys.public_keys_ptr(), _stream()))
    encrypted1.data = dst   # like the reference's resize: the ciphertext now owns a new buffer
    _after_product(context, encrypted1, encrypted2)


def multiply_and_relin_batch(context, encrypted1, encrypted2, relin_keys):
    """multiply_and_relin_inplace over lists of independent ciphertext pairs (one C-ABI call, ops interleaved over the
    engine's lanes); encrypted1[i] receives the product like the in-place form."""
    if len(encrypted1) != len(encrypted2):
        raise ValueError("batch sizes differ")
    if not encrypted1:
        return
    ci = encrypted1[0].chain_index
    for a, b in zip(encrypted1, encrypted2):
        _require_ntt(context, a)
        _require_ntt(context, b)
        if a.chain_index != ci or b.chain_index != ci:
            raise ValueError("encrypted1 and encrypted2 parameter mismatch")
        if a.is_ntt_form != b.is_ntt_form:
            raise ValueError("NTT form mismatch")
        if not _are_close(a.scale, b.scale):
            raise ValueError("scale mismatch")
        if a.size() != 2 or b.size() != 2:
            raise ValueError("destination_size must be 3")
    dst = [torch.empty_like(a.data) for a in encrypted1]
    n = len(dst)
    arr = ctypes.c_void_p * n
    check(lib.pfhe_multiply_and_relin_batch(context._h, ci, arr(*[_ptr(a.data) for a in encrypted1]),
                                            arr(*[_ptr(b.data) for b in encrypted2]), arr(*[_ptr(d) for d in dst]), n,
                                            relin_keys.public_keys_ptr(), _stream()))
    for a, b, d in zip(encrypted1, encrypted2, dst):
        a.data = d
        _after_product(context, a, b)


def apply_galois_inplace(context, encrypted, galois_elt, galois_keys):
    """apply_galois_inplace (src/evaluate.cu:1567-1630)."""
    if encrypted.size() > 2:
        raise ValueError("encrypted size must be 2")
    elts = context.parms.galois_elts
    if galois_elt not in elts:
        raise ValueError("Galois elt not present")
    key = galois_keys.get_relin_keys(elts.index(galois_elt))
    if _leveled(context) and encrypted.chain_index == 1:
        # keyswitch_inplace with is_relin = false under mul_tech hps_overq_leveled (eval_key_switch.cu:111-123, 141-146,
        # 168-174): the switched polynomial is scaled down by the levels FindLevelsToDrop allows, switched there, expanded
        drop = _levels_to_drop(context, encrypted.noise_scale_deg - 1, True, encrypted.is_asymmetric)
        if drop:
            l = encrypted.coeff_modulus_size()
            moved = torch.empty_like(encrypted.data)
            for k in range(2):
                check(lib.pfhe_apply_galois(context._h, _ptr(encrypted.data[k]), l, galois_elt, _ptr(moved[k]), _stream()))
            encrypted.data[0].copy_(moved[0])
            encrypted.data[1].zero_()
            check(lib.pfhe_keyswitch_leveled_inplace(context._h, _ptr(encrypted.data), _ptr(moved[1]), key.public_keys_ptr(),
                                                     drop, _stream()))
            return
    check(lib.pfhe_apply_galois_inplace(context._h, encrypted.chain_index, _ptr(encrypted.data), galois_elt,
                                        key.public_keys_ptr(), _stream()))


def _naf(value):
    """non-adjacent form used by rotate_internal (include/host/numth.h:17-34, src/evaluate.cu:1649)."""
    res = []
    sign = value < 0
    value = abs(value)
    i = 0
    while value:
        zi = 0
        if value & 1:
            zi = 2 - (value & 3)
        value = (value - zi) >> 1
        if zi:
            res.append((-1 if sign else 1) * zi * (1 << i))
        i += 1
    return res


def rotate_inplace(context, encrypted, step, galois_key):
    """rotate_inplace / rotate_internal (src/evaluate.cu:1633-1668)."""
    n = context.poly_degree
    elt = get_elt_from_step(step, n)
    if elt in context.parms.galois_elts:
        apply_galois_inplace(context, encrypted, elt, galois_key)
        return
    naf_steps = _naf(step)
    if len(naf_steps) == 1:
        raise ValueError("Galois key not present")
    for s in naf_steps:
        if abs(s) != (n >> 1):
            rotate_inplace(context, encrypted, s, galois_key)


def rotate_batch(context, encrypteds, steps, galois_key):
    """rotate_inplace over a list of distinct ciphertexts, one step each, in one C-ABI call (ops interleaved over the
    engine's lanes).  Every step must have its own Galois key (no NAF decomposition here)."""
    if len(encrypteds) != len(steps):
        raise ValueError("batch sizes differ")
    if not encrypteds:
        return
    elts = context.parms.galois_elts
    ci = encrypteds[0].chain_index
    keys = []
    for ct, s in zip(encrypteds, steps):
        if ct.size() > 2:
            raise ValueError("ciphertext size must be 2")
        _require_ntt(context, ct)
        if ct.chain_index != ci:
            raise ValueError("encrypteds parameter mismatch")
        e = get_elt_from_step(s, context.poly_degree)
        if e not in elts:
            raise ValueError("Galois key not present")
        keys.append(galois_key.get_relin_keys(elts.index(e)).public_keys_ptr().value)
    n = len(steps)
    check(lib.pfhe_rotate_batch(context._h, ci, (ctypes.c_void_p * n)(*[c.data.data_ptr() for c in encrypteds]),
                                (ctypes.c_int * n)(*steps), (ctypes.c_void_p * n)(*keys), n, _stream()))


def hoisting_inplace(context, ct, glk, steps):
    """hoisting_inplace (src/evaluate.cu:1670-1865)."""
    if ct.size() > 2:
        raise ValueError("ciphertext size must be 2")
    if context.scheme == scheme_type.bfv and ct.chain_index != 1:
        # the reference takes the first level's tool whatever the ciphertext's level (evaluate.cu:1688-1708)
        raise ValueError("BFV hoisting is built for the first data level")
    elts = context.parms.galois_elts
    ptrs = []
    for s in steps:
        e = get_elt_from_step(s, context.poly_degree)
        if e not in elts:
            raise RuntimeError("Galois key not present in hoisting")
        ptrs.append(glk.get_relin_keys(elts.index(e)).public_keys_ptr().value)
    arr = (ctypes.c_void_p * len(ptrs))(*ptrs)
    st = (ctypes.c_int * len(steps))(*steps)
    if _leveled(context):   # hps_overq_leveled: a key switch at the depth of the ciphertext (evaluate.cu:1690-1701)
        drop = _levels_to_drop(context, ct.noise_scale_deg - 1, True, ct.is_asymmetric)
        check(lib.pfhe_hoisting_leveled_inplace(context._h, _ptr(ct.data), st, len(steps), arr, drop, _stream()))
        return
    check(lib.pfhe_hoisting_inplace(context._h, ct.chain_index, _ptr(ct.data), st, len(steps), arr, _stream()))


def rescale_to_next(context, encrypted):
    """rescale_to_next (src/evaluate.cu:1545-1565)."""
    if context.scheme != scheme_type.ckks:
        raise ValueError("unsupported scheme")
    if encrypted.chain_index == context.size_Q:
        raise ValueError("end of modulus switching chain reached")
    l, n, size = encrypted.coeff_modulus_size(), context.poly_degree, encrypted.size()
    dst = torch.empty((size, l - 1, n), dtype=torch.int64, device=encrypted.data.device)
    check(lib.pfhe_rescale_to_next(context._h, encrypted.chain_index, _ptr(encrypted.data), size, _ptr(dst),
                                   _stream()))
    q_last = context.parms.coeff_modulus[l - 1]
    return PhantomCiphertext(None, dst, encrypted.chain_index + 1, encrypted.scale / float(q_last),
                             encrypted.is_ntt_form)


def mod_switch_to_next(context, encrypted):
    """mod_switch_to_next (src/evaluate.cu:1505-1543)."""
    if encrypted.chain_index == context.size_Q:
        raise ValueError("end of modulus switching chain reached")
    _require_ntt(context, encrypted)
    l, n, size = encrypted.coeff_modulus_size(), context.poly_degree, encrypted.size()
    dst = torch.empty((size, l - 1, n), dtype=torch.int64, device=encrypted.data.device)
    check(lib.pfhe_mod_switch_to_next(context._h, encrypted.chain_index, _ptr(encrypted.data), size, _ptr(dst),
                                      _stream()))
    out = PhantomCiphertext(None, dst, encrypted.chain_index + 1, encrypted.scale, encrypted.is_ntt_form)
    # (the reference builds the result in a fresh PhantomCiphertext: noiseScaleDeg_ and is_asymmetric_ restart at their defaults)
    if context.scheme == scheme_type.bgv:   # correction factor times q_last^-1 mod t (evaluate.cu:1420-1425)
        t = context.parms.plain_modulus
        q_last = context.parms.coeff_modulus[l - 1]
        out.correction_factor = encrypted.correction_factor * pow(q_last % t, -1, t) % t
    return out


def mod_switch_to_next_inplace(context, encrypted):
    """mod_switch_to_next_inplace (include/evaluate.cuh:148-151)"""
    encrypted.__dict__.update(mod_switch_to_next(context, encrypted).__dict__)


def rescale_to_next_inplace(context, encrypted):
    """rescale_to_next_inplace (include/evaluate.cuh:221-224)"""
    encrypted.__dict__.update(rescale_to_next(context, encrypted).__dict__)


def keyswitch_inplace(context, encrypted, c2, relin_keys, is_relin=True):
    """keyswitch_inplace (src/eval_key_switch.cu:95-182): encrypted[0..1] += key switch of the device polynomial c2 ([l][N],
    NTT form for CKKS / BGV, coefficient form for BFV) under relin_keys; is_relin only matters to hps_overq_leveled, where
    it selects the level-dropping rule"""
    _require_ntt(context, encrypted)
    if encrypted.size() < 2:
        raise ValueError("encrypted size must be at least 2")
    if _leveled(context) and encrypted.chain_index == 1:
        drop = _levels_to_drop(context, encrypted.noise_scale_deg - 1, not is_relin, encrypted.is_asymmetric)
        check(lib.pfhe_keyswitch_leveled_inplace(context._h, _ptr(encrypted.data), _ptr(c2), relin_keys.public_keys_ptr(), drop,
                                                 _stream()))
        return
    check(lib.pfhe_keyswitch_inplace(context._h, encrypted.chain_index, _ptr(encrypted.data), _ptr(c2),
                                     relin_keys.public_keys_ptr(), _stream()))


def mod_switch_to_inplace(context, encrypted, chain_index):
    """mod_switch_to_inplace (include/evaluate.cuh:169-177)"""
    if encrypted.chain_index > chain_index:
        raise ValueError("cannot switch to higher level modulus")
    while encrypted.chain_index != chain_index:
        nxt = mod_switch_to_next(context, encrypted)
        encrypted.__dict__.update(nxt.__dict__)


def mod_switch_to(context, encrypted, chain_index):
    out = encrypted.clone()
    mod_switch_to_inplace(context, out, chain_index)
    return out


# ---- linear operations and plaintext operands (src/evaluate.cu:14-338, 1106-1340) ------------------------------------------
def _are_close(a, b):
    """are_close<double> (include/host/common.h:342-345)"""
    return abs(a - b) < np.finfo(np.float64).eps * max(abs(a), abs(b), 1.0)


def balance_correction_factors(factor1, factor2, t):
    """balance_correction_factors (src/evaluate.cu:14-72): (f, e1, e2) with e1 * factor1 = e2 * factor2 = f mod t and
    |e1| + |e2| minimal over the remainders of the extended Euclidean algorithm on (t, factor2 / factor1)."""
    half = t // 2

    def bal(x):
        return x - t if x > half else x

    try:
        ratio = pow(factor1, -1, t) * factor2 % t
    except ValueError:
        raise RuntimeError("invalid correction factor1")
    e1, e2 = ratio, 1
    best = abs(bal(e1)) + abs(bal(e2))
    prev_a, prev_b, a, b = t, 0, ratio, 1
    while a != 0:
        q = prev_a // a
        prev_a, a = a, prev_a % a
        prev_b, b = b, prev_b - b * q
        a_mod, b_mod = a % t, b % t
        if a_mod != 0 and math.gcd(a_mod, t) == 1:
            cand = abs(bal(a_mod)) + abs(bal(b_mod))
            if cand < best:
                best, e1, e2 = cand, a_mod, b_mod
    return e1 * factor1 % t, e1, e2


def _check_pair(a, b):
    if a.chain_index != b.chain_index:
        raise ValueError("encrypted1 and encrypted2 parameter mismatch")
    if a.is_ntt_form != b.is_ntt_form:
        raise ValueError("NTT form mismatch")
    if not _are_close(a.scale, b.scale):
        raise ValueError("scale mismatch")
    if a.size() != b.size():
        raise ValueError("poly number mismatch")


def negate_inplace(context, encrypted):
    """negate_inplace (src/evaluate.cu:83-108)"""
    l = encrypted.coeff_modulus_size()
    for k in range(encrypted.size()):
        check(lib.pfhe_negate_rns_poly(context._h, _ptr(encrypted.data[k]), _ptr(encrypted.data[k]), l, _stream()))


def _add_sub(context, encrypted1, encrypted2, sub, negate):
    _check_pair(encrypted1, encrypted2)
    l = encrypted1.coeff_modulus_size()
    other = encrypted2.data
    if encrypted1.correction_factor != encrypted2.correction_factor:   # BGV: balance the factors first (:148-165)
        f, e1, e2 = balance_correction_factors(encrypted1.correction_factor, encrypted2.correction_factor,
                                               context.parms.plain_modulus)
        other = encrypted2.data.clone()
        check(lib.pfhe_multiply_scalar_rns_poly(context._h, _ptr(encrypted1.data), encrypted1.size(), e1, l, _stream()))
        check(lib.pfhe_multiply_scalar_rns_poly(context._h, _ptr(other), encrypted2.size(), e2, l, _stream()))
        encrypted1.correction_factor = f
    fn = lib.pfhe_sub_rns_poly if sub else lib.pfhe_add_rns_poly
    for k in range(encrypted1.size()):
        a, b = encrypted1.data[k], other[k]
        if negate:
            a, b = b, a
        check(fn(context._h, _ptr(a), _ptr(b), _ptr(encrypted1.data[k]), l, _stream()))


def add_inplace(context, encrypted1, encrypted2):
    """add_inplace (src/evaluate.cu:115-197)"""
    _add_sub(context, encrypted1, encrypted2, False, False)


def sub_inplace(context, encrypted1, encrypted2, negate=False):
    """sub_inplace (src/evaluate.cu:263-338): encrypted1 - encrypted2, or encrypted2 - encrypted1 with negate"""
    _add_sub(context, encrypted1, encrypted2, True, negate)


def add_many(context, encrypteds):
    """add_many (src/evaluate.cu:200-261) -> the sum as a new ciphertext"""
    if not encrypteds:
        raise ValueError("encrypteds cannot be empty")
    for c in encrypteds[1:]:
        try:
            _check_pair(encrypteds[0], c)
        except ValueError as e:
            raise ValueError(str(e).replace("encrypted1 and encrypted2", "encrypteds"))
    out = encrypteds[0].clone()
    for c in encrypteds[1:]:
        add_inplace(context, out, c)
    return out


def _plain_operand(context, encrypted, plain, plain_scale):
    _require_ntt(context, encrypted)
    n, l = context.poly_degree, encrypted.coeff_modulus_size()
    if context.scheme == scheme_type.ckks:
        if tuple(plain.shape) != (l, n):
            raise ValueError("encrypted and plain parameter mismatch")
    elif tuple(plain.shape) != (n,):
        raise ValueError("BFV / BGV plaintext is [N] residues mod t")
    return plain.contiguous(), (encrypted.scale if plain_scale is None else plain_scale)


def add_plain_inplace(context, encrypted, plain, plain_scale=None):
    """add_plain_inplace (src/evaluate.cu:1106-1164).  plain: device words (BFV / BGV [N] mod t, CKKS [l][N] NTT form);
    plain_scale: the plaintext's scale (the reference reads it off the PhantomPlaintext), default the ciphertext's"""
    p, ps = _plain_operand(context, encrypted, plain, plain_scale)
    if not _are_close(encrypted.scale, ps):
        raise ValueError("scale mismatch")
    check(lib.pfhe_add_plain_inplace(context._h, encrypted.chain_index, _ptr(encrypted.data), _ptr(p),
                                     encrypted.correction_factor, _stream()))


def sub_plain_inplace(context, encrypted, plain, plain_scale=None):
    """sub_plain_inplace (src/evaluate.cu:1166-1224)"""
    p, ps = _plain_operand(context, encrypted, plain, plain_scale)
    if not _are_close(encrypted.scale, ps):
        raise ValueError("scale mismatch")
    check(lib.pfhe_sub_plain_inplace(context._h, encrypted.chain_index, _ptr(encrypted.data), _ptr(p),
                                     encrypted.correction_factor, _stream()))


def multiply_plain_inplace(context, encrypted, plain, plain_scale=1.0):
    """multiply_plain_inplace (src/evaluate.cu:1226-1340); the scale becomes encrypted.scale * plain_scale"""
    p, ps = _plain_operand(context, encrypted, plain, plain_scale)
    check(lib.pfhe_multiply_plain_inplace(context._h, encrypted.chain_index, _ptr(encrypted.data), encrypted.size(), _ptr(p),
                                          _stream()))
    encrypted.scale = encrypted.scale * ps


# the copying forms the reference's Python binding exposes (python/src/binding.cu:125-165, include/evaluate.cuh)
def _copying(fn):
    def wrapped(context, encrypted, *args, **kwargs):
        out = encrypted.clone()
        fn(context, out, *args, **kwargs)
        return out
    wrapped.__name__ = fn.__name__.replace("_inplace", "")
    wrapped.__doc__ = f"copying form of {fn.__name__}"
    return wrapped


negate = _copying(negate_inplace)
add = _copying(add_inplace)
sub = _copying(sub_inplace)
add_plain = _copying(add_plain_inplace)
sub_plain = _copying(sub_plain_inplace)
multiply_plain = _copying(multiply_plain_inplace)
multiply = _copying(multiply_inplace)
multiply_and_relin = _copying(multiply_and_relin_inplace)
relinearize = _copying(relinearize_inplace)
apply_galois = _copying(apply_galois_inplace)
rotate = _copying(rotate_inplace)
hoisting = _copying(hoisting_inplace)


def nwt_2d_radix8_forward_inplace(inout, context, coeff_modulus_size, start_modulus_idx):
    """include/ntt.cuh:172-173 (tensor of [coeff_modulus_size][N] words)."""
    check(lib.pfhe_ntt_forward_inplace(context._h, _ptr(inout), coeff_modulus_size, start_modulus_idx, _stream()))


def nwt_2d_radix8_backward_inplace(inout, context, coeff_modulus_size, start_modulus_idx):
    """include/ntt.cuh:203-204"""
    check(lib.pfhe_ntt_backward_inplace(context._h, _ptr(inout), coeff_modulus_size, start_modulus_idx, _stream()))

"""Byte-compatible save / load of ciphertexts and keys (reference include/ciphertext.h:173-213, include/secretkey.h:
84-96,129-162,194-219,346-390): the streams stock Phantom writes can be read here and vice versa, so fixtures and
batches can be exchanged with an unmodified build.  Host-side only; the seed-compressed symmetric
form (save_symmetric / load_symmetric, include/ciphertext.h:216-307) is c0 plus the 64-byte seed of c1, which the loader
expands on the device.  include/phantom_b200.hpp writes the same streams from C++.

Layout of a ciphertext stream (little-endian, the reference writes raw struct members):
  size_t chain_index, size, poly_modulus_degree, coeff_modulus_size; double scale; uint64 correction_factor;
  size_t noiseScaleDeg; bool is_ntt_form; bool is_asymmetric; then size * coeff_modulus_size * N uint64 words.
A relinearisation key is size_t dnum followed by dnum public keys, each a ciphertext stream of [2][size_QP][N]; a Galois
key is size_t count followed by that many relinearisation keys; a secret key is size_t sk_max_power, N,
coeff_modulus_size and the powers' words.
"""
import struct

import numpy as np

_HDR = struct.Struct("<QQQQdQQ??")   # 58 bytes


def write_ciphertext(stream, words, chain_index, scale=1.0, correction_factor=1, noise_scale_deg=1, is_ntt_form=True,
                     is_asymmetric=False):
    w = np.ascontiguousarray(words, dtype=np.uint64)
    if w.ndim != 3:
        raise ValueError("ciphertext words must be [size][coeff_modulus_size][N]")
    size, l, n = w.shape
    stream.write(_HDR.pack(chain_index, size, n, l, float(scale), int(correction_factor), int(noise_scale_deg),
                           bool(is_ntt_form), bool(is_asymmetric)))
    stream.write(w.tobytes())


def read_ciphertext(stream):
    """-> (words [size][l][N], dict of the header fields)"""
    raw = stream.read(_HDR.size)
    if len(raw) != _HDR.size:
        raise ValueError("truncated ciphertext stream")
    ci, size, n, l, scale, cf, deg, ntt, asym = _HDR.unpack(raw)
    count = size * l * n
    body = stream.read(count * 8)
    if len(body) != count * 8:
        raise ValueError("truncated ciphertext stream")
    words = np.frombuffer(body, dtype=np.uint64).reshape(size, l, n).copy()
    return words, dict(chain_index=ci, scale=scale, correction_factor=cf, noise_scale_deg=deg, is_ntt_form=ntt,
                       is_asymmetric=asym)


def write_ciphertext_symmetric(stream, c0, seed, chain_index, scale=1.0, correction_factor=1, noise_scale_deg=1,
                               is_ntt_form=True):
    """PhantomCiphertext::save_symmetric (include/ciphertext.h:216-245): the header of save(), c0 only, then the 64-byte
    seed the second polynomial was drawn from."""
    w = np.ascontiguousarray(c0, dtype=np.uint64)
    if w.ndim != 2:
        raise ValueError("c0 must be [coeff_modulus_size][N]")
    if len(seed) != 64:
        raise ValueError("a seed is 64 bytes")
    l, n = w.shape
    stream.write(_HDR.pack(chain_index, 2, n, l, float(scale), int(correction_factor), int(noise_scale_deg), bool(is_ntt_form),
                           False))
    stream.write(w.tobytes())
    stream.write(bytes(seed))


def read_ciphertext_symmetric(stream):
    """-> (c0 [l][N], seed, dict of the header fields)   (load_symmetric, include/ciphertext.h:247-307)"""
    raw = stream.read(_HDR.size)
    if len(raw) != _HDR.size:
        raise ValueError("truncated ciphertext stream")
    ci, size, n, l, scale, cf, deg, ntt, asym = _HDR.unpack(raw)
    if asym:
        raise RuntimeError("Asymmetric ciphertext does not have seed.")
    if size != 2:
        raise RuntimeError("This method is only for 2-polynomial ciphertext.")
    body = stream.read(l * n * 8)
    seed = stream.read(64)
    if len(body) != l * n * 8 or len(seed) != 64:
        raise ValueError("truncated ciphertext stream")
    return np.frombuffer(body, dtype=np.uint64).reshape(l, n).copy(), seed, dict(
        chain_index=ci, scale=scale, correction_factor=cf, noise_scale_deg=deg, is_ntt_form=ntt, is_asymmetric=False)


def write_relin_key(stream, digits):
    """digits: dnum arrays [2][size_QP][N] (NTT form).  Header fields as generate_one_kswitch_key leaves them:
    chain_index 0, scale 1, NTT form (secretkey.cu:297-334)."""
    stream.write(struct.pack("<Q", len(digits)))
    for d in digits:
        write_ciphertext(stream, d, 0, 1.0, 1, 1, True, False)


def read_relin_key(stream):
    (dnum,) = struct.unpack("<Q", stream.read(8))
    return [read_ciphertext(stream)[0] for _ in range(dnum)]


def write_public_key(stream, pk):
    """PhantomPublicKey::save (include/secretkey.h:85-90): the ciphertext stream of pk_ -- [2][size_QP][N], chain_index 0,
    NTT form, as encrypt_zero_symmetric leaves it (secretkey.cu:380-392)."""
    write_ciphertext(stream, pk, 0, 1.0, 1, 1, True, False)


def read_public_key(stream):
    words, hdr = read_ciphertext(stream)
    if words.shape[0] != 2 or hdr["chain_index"] != 0:
        raise ValueError("not a public key stream")
    return words


def write_plaintext(stream, words, chain_index, scale=1.0):
    """PhantomPlaintext::save (include/plaintext.h:69-81): chain_index, N, coeff_modulus_size, scale, words."""
    w = np.ascontiguousarray(words, dtype=np.uint64)
    if w.ndim == 1:
        w = w.reshape(1, -1)
    if w.ndim != 2:
        raise ValueError("plaintext words must be [coeff_modulus_size][N]")
    stream.write(struct.pack("<QQQd", chain_index, w.shape[1], w.shape[0], float(scale)))
    stream.write(w.tobytes())


def read_plaintext(stream):
    """-> (words [coeff_modulus_size][N], chain_index, scale)   (PhantomPlaintext::load, include/plaintext.h:83-97)"""
    raw = stream.read(32)
    if len(raw) != 32:
        raise ValueError("truncated plaintext stream")
    ci, n, l, scale = struct.unpack("<QQQd", raw)
    body = stream.read(l * n * 8)
    if len(body) != l * n * 8:
        raise ValueError("truncated plaintext stream")
    return np.frombuffer(body, dtype=np.uint64).reshape(l, n).copy(), ci, scale


def write_galois_key(stream, keys):
    stream.write(struct.pack("<Q", len(keys)))
    for k in keys:
        write_relin_key(stream, k)


def read_galois_key(stream):
    (count,) = struct.unpack("<Q", stream.read(8))
    return [read_relin_key(stream) for _ in range(count)]


def write_secret_key(stream, powers):
    p = np.ascontiguousarray(powers, dtype=np.uint64)
    if p.ndim != 3:
        raise ValueError("secret key powers must be [sk_max_power][coeff_modulus_size][N]")
    stream.write(struct.pack("<QQQ", p.shape[0], p.shape[2], p.shape[1]))
    stream.write(p.tobytes())


def read_secret_key(stream):
    power, n, l = struct.unpack("<QQQ", stream.read(24))
    body = stream.read(power * n * l * 8)
    if len(body) != power * n * l * 8:
        raise ValueError("truncated secret key stream")
    return np.frombuffer(body, dtype=np.uint64).reshape(power, l, n).copy()

"""The names and call shapes of the reference's Python binding (``pyPhantom``, python/src/binding.cu:13-166) on top of this
engine, so that a script written against the stock binding runs with ``import pyPhantom as phantom`` unchanged (the
repository root carries a ``pyPhantom.py`` that loads this module).

The binding moves opaque C++ objects around; the only one this package does not already have is ``plaintext``
(PhantomPlaintext, include/plaintext.h): device words with the chain index and the scale the encoders, ``decrypt`` and the
plaintext operands of the evaluator read off it.  Everything else forwards to ``api.py``.
"""
import numpy as np
import torch

from . import api
from .api import (scheme_type, mul_tech_type, get_elt_from_step, get_elts_from_steps,  # noqa: F401
                  create_coeff_modulus, create_plain_modulus)

params = api.EncryptionParameters
context = api.PhantomContext
relin_key = api.PhantomRelinKey
galois_key = api.PhantomGaloisKey


class sec_level_type:
    """phantom::arith::sec_level_type (python/src/binding.cu:31-36); parameters are not checked against it here either"""
    none, tc128, tc192, tc256 = 0, 128, 192, 256


class modulus(int):
    """phantom::arith::Modulus as far as the binding exposes it: a value"""


class cuda_stream:
    """phantom::util::cuda_stream_wrapper (python/src/binding.cu:54-55): a stream object scripts may create"""

    def __init__(self):
        self.stream = torch.cuda.Stream()


class plaintext:
    """PhantomPlaintext (include/plaintext.h:10-98): BFV / BGV [N] residues mod t (chain_index 0, scale 1), CKKS [l][N] in
    NTT form at chain_index with a scale"""

    def __init__(self, data=None, chain_index=0, scale=1.0):
        self.data, self._chain_index, self._scale = data, chain_index, scale

    def chain_index(self):
        return self._chain_index

    def scale(self):
        return self._scale

    def save(self, stream):
        api.save_plaintext(stream, self.data, self._chain_index, self._scale)

    @classmethod
    def load(cls, ctx, stream):
        return cls(*api.load_plaintext(ctx, stream))


class ciphertext(api.PhantomCiphertext):
    """PhantomCiphertext as the binding exposes it: default-constructible, set_scale"""

    def set_scale(self, scale):
        self.scale = scale


def _as_ciphertext(ct):
    ct.__class__ = ciphertext
    return ct


class public_key(api.PhantomPublicKey):
    def __init__(self, ctx=None, pk=None):
        super().__init__(ctx, pk)

    def encrypt_asymmetric(self, ctx, plain):
        return _as_ciphertext(super().encrypt_asymmetric(ctx, plain.data, plain.scale()))


class secret_key(api.PhantomSecretKey):
    def __init__(self, ctx):
        super().__init__(ctx)

    def gen_publickey(self, ctx):
        return public_key(ctx, super().gen_publickey(ctx).pk)

    def encrypt_symmetric(self, ctx, plain):
        return _as_ciphertext(super().encrypt_symmetric(ctx, plain.data, plain.scale()))

    def decrypt(self, ctx, cipher):
        words = super().decrypt(ctx, cipher)
        if ctx.scheme == scheme_type.ckks:
            return plaintext(words, cipher.chain_index, cipher.scale)   # secretkey.cu:705-712
        return plaintext(words, 0, 1.0)


class batch_encoder(api.PhantomBatchEncoder):
    def encode(self, ctx, values):
        return plaintext(super().encode(ctx, np.asarray(values)), 0, 1.0)

    def decode(self, ctx, plain):
        return [int(v) for v in super().decode(ctx, plain.data)]


class ckks_encoder(api.PhantomCKKSEncoder):
    def encode_complex_vector(self, ctx, values, scale, chain_index=1):
        return plaintext(self.encode(ctx, np.asarray(values, dtype=np.complex128), scale, chain_index), chain_index, scale)

    def encode_double_vector(self, ctx, values, scale, chain_index=1):
        return plaintext(self.encode(ctx, np.asarray(values, dtype=np.float64), scale, chain_index), chain_index, scale)

    def decode_complex_vector(self, ctx, plain):
        return [complex(v) for v in self.decode(ctx, plain.data, plain.scale(), plain.chain_index())]

    def decode_double_vector(self, ctx, plain):
        return [float(v) for v in self.decode(ctx, plain.data, plain.scale(), plain.chain_index()).real]


# ---- evaluator, copying forms (python/src/binding.cu:125-165; include/evaluate.cuh) ----------------------------------------
def _check_plain_level(ctx, encrypted, plain):
    if ctx.scheme == scheme_type.ckks and plain.chain_index() != encrypted.chain_index:
        raise ValueError("encrypted and plain parameter mismatch")


def negate(ctx, encrypted):
    return _as_ciphertext(api.negate(ctx, encrypted))


def add(ctx, encrypted1, encrypted2):
    return _as_ciphertext(api.add(ctx, encrypted1, encrypted2))


def add_plain(ctx, encrypted, plain):
    _check_plain_level(ctx, encrypted, plain)
    return _as_ciphertext(api.add_plain(ctx, encrypted, plain.data, plain.scale() if ctx.scheme == scheme_type.ckks else None))


def add_many(ctx, encrypteds, destination=None):
    """add_many(context, encrypteds, destination) (src/evaluate.cu:200-261): fills `destination`; also returned"""
    total = _as_ciphertext(api.add_many(ctx, list(encrypteds)))
    if destination is not None:
        destination.__dict__.update(total.__dict__)
        return destination
    return total


def sub(ctx, encrypted1, encrypted2, negate=False):
    return _as_ciphertext(api.sub(ctx, encrypted1, encrypted2, negate))


def sub_plain(ctx, encrypted, plain):
    _check_plain_level(ctx, encrypted, plain)
    return _as_ciphertext(api.sub_plain(ctx, encrypted, plain.data, plain.scale() if ctx.scheme == scheme_type.ckks else None))


def multiply(ctx, encrypted1, encrypted2):
    return _as_ciphertext(api.multiply(ctx, encrypted1, encrypted2))


def multiply_and_relin(ctx, encrypted1, encrypted2, relin_keys):
    return _as_ciphertext(api.multiply_and_relin(ctx, encrypted1, encrypted2, relin_keys))


def multiply_plain(ctx, encrypted, plain):
    _check_plain_level(ctx, encrypted, plain)
    return _as_ciphertext(api.multiply_plain(ctx, encrypted, plain.data, plain.scale()))


def relinearize(ctx, encrypted, relin_keys):
    return _as_ciphertext(api.relinearize(ctx, encrypted, relin_keys))


def rescale_to_next(ctx, encrypted):
    return _as_ciphertext(api.rescale_to_next(ctx, encrypted))


def mod_switch_to_next(ctx, operand):
    """ciphertext (src/evaluate.cu:1505-1543) or plaintext (:1474-1503: CKKS plaintexts lose their last limb)"""
    if isinstance(operand, plaintext):
        if operand.chain_index() == ctx.size_Q:
            raise ValueError("end of modulus switching chain reached")
        nxt = operand.chain_index() + 1
        return plaintext(operand.data[:ctx.coeff_modulus_size(nxt)].clone(), nxt, operand.scale())
    return _as_ciphertext(api.mod_switch_to_next(ctx, operand))


def mod_switch_to(ctx, operand, chain_index):
    if isinstance(operand, plaintext):
        if operand.chain_index() > chain_index:
            raise ValueError("cannot switch to higher level modulus")
        while operand.chain_index() != chain_index:
            operand = mod_switch_to_next(ctx, operand)
        return operand
    return _as_ciphertext(api.mod_switch_to(ctx, operand, chain_index))


def apply_galois(ctx, encrypted, galois_elt, galois_keys):
    return _as_ciphertext(api.apply_galois(ctx, encrypted, galois_elt, galois_keys))


def rotate(ctx, encrypted, step, galois_keys):
    return _as_ciphertext(api.rotate(ctx, encrypted, step, galois_keys))


def hoisting(ctx, encrypted, galois_keys, steps):
    return _as_ciphertext(api.hoisting(ctx, encrypted, galois_keys, list(steps)))

"""phantom-fhe_b200 -- B200-native RNS polynomial-arithmetic engine behind Phantom-FHE's evaluate surface.

The compute path is the hand-written sm_100a library ``libpfhe_b200.so`` (csrc/, C-ABI in include/pfhe_b200.h).
This package is the thin host-side mirror of the reference's interface for the hot path (PhantomContext /
PhantomCiphertext / multiply_inplace / relinearize_inplace / rotate_inplace / rescale_to_next /
mod_switch_to_next, reference include/evaluate.cuh:37-245); torch is used for device memory, streams and
torch.distributed only.  There is no CPU fallback: a missing library or a missing GPU raises.
"""
from ._lib import lib, LIB_PATH, PfheError, check  # noqa: F401
from .api import (  # noqa: F401
    EncryptionParameters, PhantomContext, PhantomCiphertext, PhantomRelinKey, PhantomGaloisKey, PhantomSecretKey, PhantomPublicKey, random_bytes, save_plaintext, load_plaintext, PhantomBatchEncoder, PhantomCKKSEncoder, CoeffModulus, PlainModulus, create_coeff_modulus, create_plain_modulus,
    scheme_type, mul_tech_type, multiply_inplace, relinearize_inplace, multiply_and_relin_inplace, multiply_and_relin_batch, rotate_inplace, rotate_batch,
    apply_galois_inplace, hoisting_inplace, rescale_to_next, mod_switch_to_next, get_elt_from_step, get_elts_from_steps,
    nwt_2d_radix8_forward_inplace, nwt_2d_radix8_backward_inplace,
    negate_inplace, add_inplace, sub_inplace, add_many, add_plain_inplace, sub_plain_inplace, multiply_plain_inplace, mod_switch_to_inplace, mod_switch_to, mod_switch_to_next_inplace, rescale_to_next_inplace, keyswitch_inplace,
    negate, add, sub, add_plain, sub_plain, multiply_plain, multiply, multiply_and_relin, relinearize, apply_galois, rotate, hoisting, balance_correction_factors,
)

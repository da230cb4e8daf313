"""ctypes binding of libpfhe_b200.so (include/pfhe_b200.h).  Fails loudly if the CUDA library is missing."""
import ctypes
import os

_here = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PFHE_B200_LIB") or os.path.join(_here, "libpfhe_b200.so")   # override: A/B builds

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} not found: build the sm_100a extension first (python -c 'import __graft_entry__ as g; g.build()' "
        "or make -C phantom-fhe_b200/csrc).  This package has no CPU fallback.")

lib = ctypes.CDLL(LIB_PATH)

u64p = ctypes.POINTER(ctypes.c_uint64)
u32p = ctypes.POINTER(ctypes.c_uint32)
i32p = ctypes.POINTER(ctypes.c_int)
vp = ctypes.c_void_p
sz = ctypes.c_size_t

_sigs = {
    "pfhe_last_error": (ctypes.c_char_p, []),
    "pfhe_create_primes": (ctypes.c_int, [ctypes.c_uint64, i32p, ctypes.c_int, u64p]),
    "pfhe_engine_create": (ctypes.c_int, [ctypes.POINTER(vp), ctypes.c_int, ctypes.c_uint64, u64p, ctypes.c_int,
                                          ctypes.c_int, ctypes.c_uint64, u32p, ctypes.c_int]),
    "pfhe_engine_destroy": (None, [vp]),
    "pfhe_engine_set_mul_tech": (ctypes.c_int, [vp, ctypes.c_int]),
    "pfhe_poly_degree": (ctypes.c_uint64, [vp]),
    "pfhe_size_QP": (ctypes.c_int, [vp]),
    "pfhe_size_P": (ctypes.c_int, [vp]),
    "pfhe_dnum": (ctypes.c_int, [vp, sz]),
    "pfhe_galois_elts": (ctypes.c_int, [vp, u32p, ctypes.c_int]),
    "pfhe_galois_elt_from_step": (ctypes.c_int, [ctypes.c_int, ctypes.c_uint64, u32p]),
    "pfhe_ntt_forward_inplace": (ctypes.c_int, [vp, vp, sz, sz, vp]),
    "pfhe_ntt_backward_inplace": (ctypes.c_int, [vp, vp, sz, sz, vp]),
    "pfhe_ntt_forward_inplace_batch": (ctypes.c_int, [vp, vp, sz, sz, sz, vp]),
    "pfhe_ntt_backward_inplace_batch": (ctypes.c_int, [vp, vp, sz, sz, sz, vp]),
    "pfhe_ntt_backward": (ctypes.c_int, [vp, vp, vp, sz, sz, vp]),
    "pfhe_ntt_forward_inplace_include_special_mod": (ctypes.c_int, [vp, vp, sz, sz, sz, sz, vp]),
    "pfhe_ntt_backward_inplace_include_special_mod": (ctypes.c_int, [vp, vp, sz, sz, sz, sz, vp]),
    "pfhe_table_size": (ctypes.c_int, [vp, ctypes.c_int]),
    "pfhe_table_modulus": (ctypes.c_uint64, [vp, ctypes.c_int, sz]),
    "pfhe_nwt_2d_radix8_forward_inplace": (ctypes.c_int, [vp, ctypes.c_int, vp, sz, sz, vp]),
    "pfhe_nwt_2d_radix8_backward_inplace": (ctypes.c_int, [vp, ctypes.c_int, vp, sz, sz, vp]),
    "pfhe_nwt_2d_radix8_backward": (ctypes.c_int, [vp, ctypes.c_int, vp, vp, sz, sz, vp]),
    "pfhe_nwt_2d_radix8_forward_inplace_fuse_moddown": (ctypes.c_int, [vp, vp, vp, vp, vp, vp, sz, sz, vp]),
    "pfhe_nwt_2d_radix8_forward_inplace_include_temp_mod": (ctypes.c_int, [vp, ctypes.c_int, vp, sz, sz, sz, vp]),
    "pfhe_nwt_2d_radix8_forward_inplace_include_special_mod_exclude_range": (ctypes.c_int, [vp, vp, sz, sz, sz, sz, sz, sz, vp]),
    "pfhe_nwt_2d_radix8_forward_modup_fuse": (ctypes.c_int, [vp, vp, vp, sz, sz, sz, vp]),
    "pfhe_nwt_2d_radix8_backward_scale": (ctypes.c_int, [vp, ctypes.c_int, vp, vp, sz, sz, vp, vp, vp]),
    "pfhe_nwt_2d_radix8_backward_inplace_scale": (ctypes.c_int, [vp, ctypes.c_int, vp, sz, sz, vp, vp, vp]),
    "pfhe_nwt_2d_radix8_backward_inplace_include_temp_mod_scale": (ctypes.c_int, [vp, ctypes.c_int, vp, sz, sz, sz, vp, vp, vp]),
    "pfhe_bconv": (ctypes.c_int, [vp, ctypes.c_int, u32p, ctypes.c_int, u32p, ctypes.c_int, vp, vp, vp]),
    "pfhe_moddown": (ctypes.c_int, [vp, sz, vp, vp, vp]),
    "pfhe_divide_and_round_q_last": (ctypes.c_int, [vp, sz, vp, sz, vp, vp]),
    "pfhe_divide_and_round_q_last_ntt": (ctypes.c_int, [vp, sz, vp, sz, vp, vp]),
    "pfhe_mod_t_and_divide_q_last_ntt": (ctypes.c_int, [vp, sz, vp, sz, vp, vp]),
    "pfhe_add_to_ct": (ctypes.c_int, [vp, vp, vp, sz, vp]),
    "pfhe_tensor_prod_2x2": (ctypes.c_int, [vp, vp, vp, vp, sz, vp]),
    "pfhe_tensor_square_2x2": (ctypes.c_int, [vp, vp, vp, sz, vp]),
    "pfhe_add_rns_poly": (ctypes.c_int, [vp, vp, vp, vp, sz, vp]),
    "pfhe_sub_rns_poly": (ctypes.c_int, [vp, vp, vp, vp, sz, vp]),
    "pfhe_multiply_rns_poly": (ctypes.c_int, [vp, vp, vp, vp, sz, vp]),
    "pfhe_negate_rns_poly": (ctypes.c_int, [vp, vp, vp, sz, vp]),
    "pfhe_modup": (ctypes.c_int, [vp, sz, vp, vp, vp]),
    "pfhe_key_switch_inner_prod": (ctypes.c_int, [vp, sz, vp, vp, vp, vp]),
    "pfhe_moddown_from_ntt": (ctypes.c_int, [vp, sz, vp, vp, vp]),
    "pfhe_keyswitch_inplace": (ctypes.c_int, [vp, sz, vp, vp, vp, vp]),
    "pfhe_multiply_and_relin_inplace": (ctypes.c_int, [vp, sz, vp, vp, vp, vp]),
    "pfhe_multiply_and_relin": (ctypes.c_int, [vp, sz, vp, vp, vp, vp, vp]),
    "pfhe_multiply": (ctypes.c_int, [vp, sz, vp, vp, vp, vp]),
    "pfhe_relinearize_inplace": (ctypes.c_int, [vp, sz, vp, vp, vp]),
    "pfhe_apply_galois_inplace": (ctypes.c_int, [vp, sz, vp, ctypes.c_uint32, vp, vp]),
    "pfhe_rotate_batch": (ctypes.c_int, [vp, sz, vp, i32p, vp, sz, vp]),
    "pfhe_rotate_inplace": (ctypes.c_int, [vp, sz, vp, ctypes.c_int, vp, vp]),
    "pfhe_hoisting_inplace": (ctypes.c_int, [vp, sz, vp, i32p, sz, vp, vp]),
    "pfhe_hoisting_leveled_inplace": (ctypes.c_int, [vp, vp, i32p, sz, vp, ctypes.c_int, vp]),
    "pfhe_rescale_to_next": (ctypes.c_int, [vp, sz, vp, sz, vp, vp]),
    "pfhe_mod_switch_to_next": (ctypes.c_int, [vp, sz, vp, sz, vp, vp]),
    "pfhe_ckks_encode": (ctypes.c_int, [vp, sz, vp, sz, ctypes.c_double, vp, vp]),
    "pfhe_ckks_decode": (ctypes.c_int, [vp, sz, vp, ctypes.c_double, vp, vp]),
    "pfhe_sample_poly": (ctypes.c_int, [vp, ctypes.c_int, sz, ctypes.c_char_p, vp, vp]),
    "pfhe_gen_secretkey": (ctypes.c_int, [vp, ctypes.c_char_p, vp, vp]),
    "pfhe_encrypt_zero_symmetric": (ctypes.c_int, [vp, sz, vp, ctypes.c_char_p, ctypes.c_char_p, vp, vp]),
    "pfhe_encrypt_zero_asymmetric": (ctypes.c_int, [vp, sz, vp, ctypes.c_char_p, ctypes.c_char_p, vp, vp]),
    "pfhe_gen_kswitch_key": (ctypes.c_int, [vp, vp, vp, ctypes.c_char_p, vp, vp]),
    "pfhe_galois_secret_key": (ctypes.c_int, [vp, vp, ctypes.c_uint32, vp, vp]),
    "pfhe_apply_galois_ntt": (ctypes.c_int, [vp, vp, sz, ctypes.c_uint32, vp, vp]),
    "pfhe_apply_galois": (ctypes.c_int, [vp, vp, sz, ctypes.c_uint32, vp, vp]),
    "pfhe_encrypt_add_plain": (ctypes.c_int, [vp, sz, vp, vp, vp]),
    "pfhe_add_plain_inplace": (ctypes.c_int, [vp, sz, vp, vp, ctypes.c_uint64, vp]),
    "pfhe_sub_plain_inplace": (ctypes.c_int, [vp, sz, vp, vp, ctypes.c_uint64, vp]),
    "pfhe_multiply_plain_inplace": (ctypes.c_int, [vp, sz, vp, sz, vp, vp]),
    "pfhe_multiply_scalar_rns_poly": (ctypes.c_int, [vp, vp, sz, ctypes.c_uint64, sz, vp]),
    "pfhe_batch_encode": (ctypes.c_int, [vp, vp, sz, vp, vp]),
    "pfhe_batch_decode": (ctypes.c_int, [vp, vp, vp, vp]),
    "pfhe_decrypt": (ctypes.c_int, [vp, sz, vp, sz, vp, ctypes.c_uint64, vp, vp]),
    "pfhe_find_levels_to_drop": (ctypes.c_int, [vp, sz, ctypes.c_int, ctypes.c_int, i32p]),
    "pfhe_multiply_leveled": (ctypes.c_int, [vp, vp, vp, vp, ctypes.c_int, vp]),
    "pfhe_multiply_and_relin_leveled": (ctypes.c_int, [vp, vp, vp, vp, vp, ctypes.c_int, vp]),
    "pfhe_keyswitch_leveled_inplace": (ctypes.c_int, [vp, vp, vp, vp, ctypes.c_int, vp]),
    "pfhe_fnwt_1d": (ctypes.c_int, [vp, vp, vp, vp, sz, sz, sz, vp]),
    "pfhe_inwt_1d": (ctypes.c_int, [vp, vp, vp, vp, vp, vp, sz, sz, sz, vp]),
    "pfhe_multiply_sizes": (ctypes.c_int, [vp, sz, vp, sz, vp, sz, vp, vp]),
    "pfhe_multiply_and_relin_batch": (ctypes.c_int, [vp, sz, vp, vp, vp, sz, vp, vp]),
    "pfhe_engine_set_lanes": (ctypes.c_int, [vp, ctypes.c_int]),
    "pfhe_engine_lanes": (ctypes.c_int, [vp]),
    "pfhe_multiply_and_relin_host": (ctypes.c_int, [vp, sz, vp, vp, vp, vp, vp]),
    "pfhe_multiply_and_relin_host_batch": (ctypes.c_int, [vp, sz, vp, vp, vp, sz, vp, vp]),
    "pfhe_rotate_host": (ctypes.c_int, [vp, sz, vp, ctypes.c_int, vp, vp, vp]),
    "pfhe_rescale_host": (ctypes.c_int, [vp, sz, vp, sz, vp, vp]),
    "pfhe_ntt_forward_host": (ctypes.c_int, [vp, vp, vp, sz, sz, vp]),
    "pfhe_launch_count": (ctypes.c_uint64, [vp]),
    "pfhe_enable_peer_access": (ctypes.c_int, [ctypes.c_int]),
    "pfhe_ipc_export": (ctypes.c_int, [vp, ctypes.c_char_p, u64p]),
    "pfhe_ipc_open": (ctypes.c_int, [ctypes.c_char_p, ctypes.c_uint64, ctypes.POINTER(vp)]),
    "pfhe_ipc_close": (ctypes.c_int, [vp, ctypes.c_uint64]),
}
for _name, (_res, _args) in _sigs.items():
    _f = getattr(lib, _name)   # AttributeError if the library does not export a declared symbol
    _f.restype = _res
    _f.argtypes = _args

PFHE_OK, PFHE_ERR_INVALID_ARGUMENT, PFHE_ERR_LOGIC, PFHE_ERR_CUDA, PFHE_ERR_UNSUPPORTED = range(5)


class PfheError(RuntimeError):
    """CUDA failure inside the engine (reference: std::runtime_error("CUDA Runtime Error"))."""


def check(status):
    """Re-raise engine status codes as the exception class the reference throws for the same condition:
    std::invalid_argument -> ValueError, std::logic_error -> RuntimeError."""
    if status == PFHE_OK:
        return
    msg = lib.pfhe_last_error().decode()
    if status == PFHE_ERR_INVALID_ARGUMENT:
        raise ValueError(msg)
    if status in (PFHE_ERR_LOGIC, PFHE_ERR_UNSUPPORTED):
        raise RuntimeError(msg)
    raise PfheError(msg)

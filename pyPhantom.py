"""``import pyPhantom as phantom``: the module name of the reference's Python binding (python/setup.py), served by
phantom-fhe_b200/pyphantom.py on top of the B200 engine."""
import phantom_fhe_b200  # noqa: F401  (registers the package under an importable name)
from phantom_fhe_b200.pyphantom import *  # noqa: F401,F403
from phantom_fhe_b200.pyphantom import (params, context, secret_key, public_key, relin_key, galois_key, batch_encoder,  # noqa: F401
                                        ckks_encoder, plaintext, ciphertext, modulus, cuda_stream, sec_level_type)

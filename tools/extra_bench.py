"""BASELINE.json configs[2] (BFV HMult+Relin, N=2^14, log QP = 438, t = 65537) and configs[3] (CKKS rotate, N=2^16, L=16,
32 rotation steps) on one GPU, device resident, CUDA events.  Called by bench.py for the `extra` key of its JSON line;
product API only (synthetic uniform residues and synthetic keys made on the device)."""
import ctypes


def _fill(torch, dev, view, moduli, n, gen):
    flat = view.reshape(-1, len(moduli), n)
    for j, q in enumerate(moduli):
        flat[:, j, :] = torch.randint(0, int(q), (flat.shape[0], n), generator=gen, device=dev, dtype=torch.int64)


def _key(pf, torch, dev, ctx, primes, n, gen):
    digits = [torch.empty((2, len(primes), n), dtype=torch.int64, device=dev) for _ in range(ctx.dnum(1))]
    for d in digits:
        _fill(torch, dev, d, primes, n, gen)
    return pf.PhantomRelinKey.from_device(ctx, digits)


def _time(torch, fn, reps, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps


def bfv_c3(pf, lib, check, torch, dev, reps=100):
    n = 16384
    primes = pf.CoeffModulus.Create(n, [54] * 7 + [60])
    out = {"workload": "BFV HMult+Relin, N=2^14, primes {54x7, 60}, special_modulus_size 1 (log QP = 438), t = 65537 "
                       "(benchmark/bfv_bench.cu:305-318), one op at a time, device resident", "unit": "us per op"}
    gen = torch.Generator(device=dev)
    gen.manual_seed(7)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for tech in (pf.mul_tech_type.behz, pf.mul_tech_type.hps, pf.mul_tech_type.hps_overq):
        parms = pf.EncryptionParameters(pf.scheme_type.bfv)
        parms.set_poly_modulus_degree(n)
        parms.set_coeff_modulus(primes)
        parms.set_special_modulus_size(1)
        parms.set_plain_modulus(65537)
        parms.set_mul_tech(tech)
        ctx = pf.PhantomContext(parms)
        l = len(primes) - 1
        a = torch.empty((2, l, n), dtype=torch.int64, device=dev)
        b = torch.empty_like(a)
        o = torch.empty_like(a)
        _fill(torch, dev, a, primes[:l], n, gen)
        _fill(torch, dev, b, primes[:l], n, gen)
        rlk = _key(pf, torch, dev, ctx, primes, n, gen)
        us = _time(torch, lambda: check(lib.pfhe_multiply_and_relin(ctx._h, 1, a.data_ptr(), b.data_ptr(), o.data_ptr(),
                                                                    rlk.public_keys_ptr(), st)), reps)
        out[tech.name] = us
        del ctx
    return out


def rotate_c4(pf, lib, check, torch, dev, store=None):
    n = 65536
    primes = pf.CoeffModulus.Create(n, [60] + [40] * 15 + [60] * 4)
    steps = list(range(1, 33))
    parms = pf.EncryptionParameters(pf.scheme_type.ckks)
    parms.set_poly_modulus_degree(n)
    parms.set_coeff_modulus(primes)
    parms.set_special_modulus_size(4)
    parms.set_galois_elts(pf.get_elts_from_steps(steps, n))
    ctx = pf.PhantomContext(parms)
    l = 16
    gen = torch.Generator(device=dev)
    gen.manual_seed(9)
    keys = [_key(pf, torch, dev, ctx, primes, n, gen) for _ in steps]   # 32 x 80 MiB
    glk = pf.PhantomGaloisKey.__new__(pf.PhantomGaloisKey)
    glk.relin_keys = keys
    src = torch.empty((2, l, n), dtype=torch.int64, device=dev)
    _fill(torch, dev, src, primes[:l], n, gen)
    cts = [pf.PhantomCiphertext(ctx, src.clone()) for _ in steps]

    def refill():
        for c in cts:
            c.data.copy_(src)

    def timed(fn, reps=5):
        best = 1e30
        for _ in range(reps):
            refill()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3)
        return best

    def one_by_one():
        for c, s in zip(cts, steps):
            pf.rotate_inplace(ctx, c, s, glk)

    out = {"workload": "CKKS rotate_inplace, N=2^16, L=16, alpha=4: steps 1..32, each on a fresh copy of one ciphertext "
                       "(keyswitch_bench.cu pattern), device resident; best of 5", "unit": "us for the 32 rotations"}
    out["one_at_a_time"] = timed(one_by_one)
    out["rotate_batch"] = timed(lambda: pf.rotate_batch(ctx, cts, steps, glk))
    h = pf.PhantomCiphertext(ctx, src.clone())
    out["hoisting_inplace_sum"] = timed(lambda: pf.hoisting_inplace(ctx, h, glk, steps))
    return out


def run(pf, lib, check, torch, dev):
    return {"c3_bfv": bfv_c3(pf, lib, check, torch, dev), "c4_rotate32": rotate_c4(pf, lib, check, torch, dev)}

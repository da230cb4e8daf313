// An application written against the reference's class and function names (phantom.h), running on the C++ mirror
// include/phantom_b200.hpp: key generation, encoding, public-key and symmetric encryption, multiply + relinearize (fused
// and in two steps), rotation, addition, level switching, rescaling, decryption, decoding -- BFV (mul_tech
// hps_overq_leveled), BGV and CKKS.  Prints OK and exits 0 when every decrypted result is the expected one.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "phantom_b200.hpp"

using namespace phantom_b200;

static int failures = 0;
// ring degree: 2^13 unless PFHE_DEMO_LOGN says otherwise (the CPU run against the oracle uses a smaller ring)
static size_t ring_degree() {
    const char *v = std::getenv("PFHE_DEMO_LOGN");
    return (size_t) 1 << (v ? std::atoi(v) : 13);
}
static void expect(bool ok, const std::string &what) {
    std::printf("%s %s\n", ok ? "ok  " : "FAIL", what.c_str());
    if (!ok) failures++;
}

// every object through its stream and back (the reference's save / load); with PFHE_DEMO_DUMP=<dir> the streams are also
// written to files there (tests/test_host_logic.py reads them with the Python mirror's serial.py)
template<class T>
static std::string saved(const T &object) {
    std::ostringstream out;
    object.save(out);
    return out.str();
}
static void dump(const std::string &name, const std::string &bytes) {
    if (const char *dir = std::getenv("PFHE_DEMO_DUMP")) {
        std::ofstream f(std::string(dir) + "/" + name, std::ios::binary);
        f.write(bytes.data(), (std::streamsize) bytes.size());
    }
}
template<class T>
static void reload(T &object, const std::string &name) {
    const std::string bytes = saved(object);
    dump(name, bytes);
    std::istringstream in(bytes);
    T fresh;
    fresh.load(in);
    expect(saved(fresh) == bytes, name + ": save, load, save gives the same stream");
    object = std::move(fresh);
}

static void integer_scheme(scheme_type scheme) {
    const size_t n = ring_degree();
    const std::string name = scheme == scheme_type::bfv ? "bfv" : "bgv";
    EncryptionParameters parms(scheme);
    parms.set_poly_modulus_degree(n);
    parms.set_coeff_modulus(CoeffModulus::Create(n, {50, 50, 50, 50, 60, 60}));
    parms.set_special_modulus_size(2);
    parms.set_plain_modulus(PlainModulus::Batching(n, 20));
    parms.set_galois_elts(get_elts_from_steps({1}, n));
    if (scheme == scheme_type::bfv) parms.set_mul_tech(mul_tech_type::hps_overq_leveled);
    PhantomContext context(parms);
    const uint64_t t = parms.plain_modulus();

    PhantomSecretKey secret_key(context);
    PhantomPublicKey public_key = secret_key.gen_publickey(context);
    PhantomRelinKey relin_keys = secret_key.gen_relinkey(context);
    PhantomGaloisKey galois_keys = secret_key.create_galois_keys(context);
    PhantomBatchEncoder encoder(context);
    reload(secret_key, name + "_secret_key.bin");   // from here on everything runs on keys that went through their streams
    reload(public_key, name + "_public_key.bin");
    reload(relin_keys, name + "_relin_key.bin");
    reload(galois_keys, name + "_galois_key.bin");

    std::vector<uint64_t> msg(n), sq(n), rot(n);
    for (size_t i = 0; i < n; i++) msg[i] = (i * 7 + 3) % 1000, sq[i] = msg[i] * msg[i] % t;
    const size_t half = n / 2;
    for (size_t i = 0; i < half; i++) rot[i] = sq[(i + 1) % half], rot[half + i] = sq[half + (i + 1) % half];

    auto decrypted = [&](const PhantomCiphertext &ct) {
        PhantomPlaintext pt;
        secret_key.decrypt(context, ct, pt);
        auto v = encoder.decode(context, pt);
        for (auto &x : v) x %= t;
        return v;
    };

    PhantomPlaintext plain;
    encoder.encode(context, msg, plain);
    PhantomCiphertext asym, sym;
    public_key.encrypt_asymmetric(context, plain, asym);
    secret_key.encrypt_symmetric(context, plain, sym);
    expect(decrypted(asym) == msg, name + ": public-key encryption round trip");
    expect(decrypted(sym) == msg, name + ": symmetric encryption round trip");
    reload(plain, name + "_plaintext.bin");
    reload(asym, name + "_ciphertext.bin");
    {   // seed-compressed form: c0 and the seed of c1; the loader draws c1 again
        std::ostringstream out;
        sym.save_symmetric(out);
        dump(name + "_ciphertext_symmetric.bin", out.str());
        std::istringstream in(out.str());
        PhantomCiphertext expanded;
        expanded.load_symmetric(context, in);
        expect(out.str().size() == 58 + sym.coeff_modulus_size() * n * 8 + 64 && saved(expanded) == saved(sym),
               name + ": save_symmetric / load_symmetric rebuild the ciphertext");
        bool refused = false;
        try {
            std::ostringstream no;
            asym.save_symmetric(no);
        } catch (const std::runtime_error &) { refused = true; }
        expect(refused, name + ": a public-key ciphertext has no seed to save");
    }

    PhantomCiphertext fused = asym;
    multiply_and_relin_inplace(context, fused, asym, relin_keys);
    expect(fused.size() == 2 && decrypted(fused) == sq, name + ": multiply_and_relin_inplace");
    PhantomCiphertext two_step = sym;
    multiply_inplace(context, two_step, sym);
    expect(two_step.size() == 3 && decrypted(two_step) == sq, name + ": multiply_inplace (three polynomials)");
    relinearize_inplace(context, two_step, relin_keys);
    expect(two_step.size() == 2 && decrypted(two_step) == sq, name + ": relinearize_inplace");

    {   // hoisting over {1, 1}: twice the rotation by one step (one shared mod-up; BFV with the coefficient-form ends)
        PhantomCiphertext hoisted = fused;
        hoisting_inplace(context, hoisted, galois_keys, {1, 1});
        std::vector<uint64_t> twice_rot(n);
        for (size_t i = 0; i < n; i++) twice_rot[i] = 2 * rot[i] % t;
        expect(decrypted(hoisted) == twice_rot, name + ": hoisting_inplace");
    }
    rotate_inplace(context, fused, 1, galois_keys);
    expect(decrypted(fused) == rot, name + ": rotate_inplace by one step");

    PhantomCiphertext sum = sym;
    add_inplace(context, sum, asym);
    std::vector<uint64_t> twice(n);
    for (size_t i = 0; i < n; i++) twice[i] = 2 * msg[i] % t;
    expect(decrypted(sum) == twice, name + ": add_inplace");
    add_plain_inplace(context, sum, plain);
    sub_inplace(context, sum, asym);
    expect(decrypted(sum) == twice, name + ": add_plain_inplace, sub_inplace");
    multiply_plain_inplace(context, sym, plain);
    expect(decrypted(sym) == sq, name + ": multiply_plain_inplace");
    PhantomCiphertext lower = mod_switch_to_next(context, two_step);
    expect(lower.chain_index() == 2 && lower.coeff_modulus_size() == 3 && decrypted(lower) == sq, name + ": mod_switch_to_next");
    if (scheme == scheme_type::bgv) {
        // operands whose correction factors differ: x^2 one level down carries q_last^-2, `lower` carries q_last^-1; the
        // sum balances them (balance_correction_factors) and still decrypts to the sum of the two messages
        PhantomCiphertext fresh;
        secret_key.encrypt_symmetric(context, plain, fresh);
        PhantomCiphertext x = mod_switch_to_next(context, fresh);
        PhantomCiphertext xx = x;
        multiply_and_relin_inplace(context, xx, x, relin_keys);
        expect(xx.correction_factor() != lower.correction_factor() && decrypted(xx) == sq, name + ": product one level down");
        add_inplace(context, xx, lower);
        std::vector<uint64_t> twice_sq(n);
        for (size_t i = 0; i < n; i++) twice_sq[i] = 2 * sq[i] % t;
        expect(decrypted(xx) == twice_sq, name + ": add_inplace balances different correction factors");
        sub_inplace(context, xx, lower, true);   // lower - xx = -sq
        std::vector<uint64_t> minus_sq(n);
        for (size_t i = 0; i < n; i++) minus_sq[i] = (t - sq[i]) % t;
        expect(decrypted(xx) == minus_sq, name + ": sub_inplace with negate after balancing");
    }
    bool threw = false;
    try {
        add_inplace(context, lower, two_step);
    } catch (const std::invalid_argument &) { threw = true; }
    expect(threw, name + ": level mismatch is refused");
}

static void ckks() {
    const size_t n = ring_degree();
    const double scale = 1099511627776.0;   // 2^40
    EncryptionParameters parms(scheme_type::ckks);
    parms.set_poly_modulus_degree(n);
    parms.set_coeff_modulus(CoeffModulus::Create(n, {60, 40, 40, 60}));
    parms.set_special_modulus_size(1);
    parms.set_galois_elts(get_elts_from_steps({1, 2, 4, -1}, n));
    PhantomContext context(parms);
    PhantomSecretKey secret_key(context);
    PhantomPublicKey public_key = secret_key.gen_publickey(context);
    PhantomRelinKey relin_keys = secret_key.gen_relinkey(context);
    PhantomGaloisKey galois_keys = secret_key.create_galois_keys(context);
    PhantomCKKSEncoder encoder(context);
    const size_t slots = encoder.slot_count();
    std::vector<double> msg(slots);
    for (size_t i = 0; i < slots; i++) msg[i] = 0.001 * (double) (i % 1000) - 0.3;

    auto max_error = [&](const PhantomCiphertext &ct, const std::vector<double> &want) {
        PhantomPlaintext pt;
        secret_key.decrypt(context, ct, pt);
        std::vector<double> got;
        encoder.decode(context, pt, got);
        double worst = 0;
        for (size_t i = 0; i < slots; i++) worst = std::max(worst, std::fabs(got[i] - want[i]));
        return worst;
    };

    PhantomPlaintext plain;
    encoder.encode(context, msg, scale, plain);
    PhantomCiphertext ct;
    public_key.encrypt_asymmetric(context, plain, ct);
    expect(max_error(ct, msg) < 1e-6, "ckks: encode, encrypt, decrypt, decode");
    reload(plain, "ckks_plaintext.bin");
    reload(ct, "ckks_ciphertext.bin");
    reload(relin_keys, "ckks_relin_key.bin");
    expect(plain.scale() == scale && plain.chain_index() == 1 && ct.scale() == scale, "ckks: level and scale survive the streams");
    PhantomCiphertext prod = ct;
    multiply_and_relin_inplace(context, prod, ct, relin_keys);
    PhantomCiphertext rescaled = rescale_to_next(context, prod);
    std::vector<double> sq(slots), rot(slots);
    for (size_t i = 0; i < slots; i++) sq[i] = msg[i] * msg[i];
    for (size_t i = 0; i < slots; i++) rot[i] = sq[(i + 2) % slots];
    expect(rescaled.chain_index() == 2 && max_error(rescaled, sq) < 1e-5, "ckks: multiply_and_relin_inplace, rescale_to_next");
    PhantomCiphertext by_three = rescaled;   // no key for step 3: composed from its non-adjacent form, -1 then +4
    rotate_inplace(context, by_three, 3, galois_keys);
    std::vector<double> rot3(slots);
    for (size_t i = 0; i < slots; i++) rot3[i] = sq[(i + 3) % slots];
    expect(max_error(by_three, rot3) < 1e-5, "ckks: rotate_inplace by three steps through the NAF recursion");
    PhantomCiphertext hoisted = rescaled;   // the rotations by 1, 2 and 4 summed with one shared mod-up
    hoisting_inplace(context, hoisted, galois_keys, {1, 2, 4});
    std::vector<double> rot_sum(slots);
    for (size_t i = 0; i < slots; i++) rot_sum[i] = sq[(i + 1) % slots] + sq[(i + 2) % slots] + sq[(i + 4) % slots];
    expect(max_error(hoisted, rot_sum) < 1e-4, "ckks: hoisting_inplace over three steps");
    rotate_inplace(context, rescaled, 2, galois_keys);
    expect(max_error(rescaled, rot) < 1e-5, "ckks: rotate_inplace by two steps");
    PhantomCiphertext sym;
    secret_key.encrypt_symmetric(context, plain, sym);
    add_inplace(context, sym, ct);
    std::vector<double> twice(slots);
    for (size_t i = 0; i < slots; i++) twice[i] = 2 * msg[i];
    expect(max_error(sym, twice) < 1e-6, "ckks: encrypt_symmetric, add_inplace");
}

int main() {
    try {
        integer_scheme(scheme_type::bfv);
        integer_scheme(scheme_type::bgv);
        ckks();
    } catch (const std::exception &e) {
        std::printf("FAIL exception: %s\n", e.what());
        return 2;
    }
    std::printf(failures ? "FAILED %d\n" : "OK\n", failures);
    return failures ? 1 : 0;
}

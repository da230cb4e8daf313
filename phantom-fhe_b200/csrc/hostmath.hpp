// hostmath.hpp -- host-side number theory used to build the engine's constant tables.
//
// Produces the same *values* as the reference's host code (include/host/*, src/host/*): NTT-friendly prime
// chains (CoeffModulus::Create, src/host/modulus.cu:79-110), the minimal primitive 2N-th root
// (try_minimal_primitive_root, src/host/numth.cu:309-331), Shoup companions
// (include/host/uintarithsmallmod.h:119-124), Barrett ratios (src/host/modulus.cu:28-41), base-conversion
// matrices (src/host/rns.cu:438-497).  Written from the mathematical definitions; header-only.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <vector>

namespace pfhe::host {

using u64 = unsigned long long;
using u128 = unsigned __int128;

inline u64 mulmod(u64 a, u64 b, u64 q) { return (u64) ((u128) a * b % q); }

inline u64 powmod(u64 a, u64 e, u64 q) {
    u64 r = 1 % q;
    a %= q;
    for (; e; e >>= 1) {
        if (e & 1) r = mulmod(r, a, q);
        a = mulmod(a, a, q);
    }
    return r;
}

// modular inverse by the extended Euclidean algorithm; throws if gcd(a, q) != 1
inline u64 invmod(u64 a, u64 q) {
    __int128 x0 = 0, x1 = 1;
    u64 r0 = q, r1 = a % q;
    while (r1 != 0) {
        u64 k = r0 / r1;
        __int128 x2 = x0 - (__int128) k * x1;
        x0 = x1, x1 = x2;
        u64 r2 = r0 - k * r1;
        r0 = r1, r1 = r2;
    }
    if (r0 != 1) throw std::invalid_argument("invmod: operand is not invertible");
    return (u64) (x0 < 0 ? x0 + q : x0);
}

inline u64 shoup(u64 w, u64 q) { return (u64) (((u128) w << 64) / q); }

struct BarrettRatio {
    u64 lo, hi;
};
// floor(2^128 / q) for q > 1
inline BarrettRatio barrett_ratio(u64 q) {
    u128 top = ((u128) 1 << 64);          // 2^64 = hi1 * q + r1
    u64 hi = (u64) (top / q);
    u64 r1 = (u64) (top % q);
    u128 rest = ((u128) r1 << 64);        // r1 * 2^64 = lo * q + r0
    u64 lo = (u64) (rest / q);
    return {lo, hi};
}

inline bool is_prime(u64 v) {
    if (v < 2) return false;
    for (u64 p : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
        if (v == p) return true;
        if (v % p == 0) return false;
    }
    u64 d = v - 1;
    int s = 0;
    while ((d & 1) == 0) d >>= 1, ++s;
    for (u64 a : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
        u64 x = powmod(a, d, v);
        if (x == 1 || x == v - 1) continue;
        bool witness = true;
        for (int i = 1; i < s && witness; ++i) {
            x = mulmod(x, x, v);
            if (x == v - 1) witness = false;
        }
        if (witness) return false;
    }
    return true;
}

// Prime chain for the given bit sizes: for every distinct size the primes congruent to 1 mod 2N are taken
// downwards from 2^bits; among equal sizes, later positions receive larger primes.
inline std::vector<u64> create_primes(u64 n, const std::vector<int> &bit_sizes) {
    std::vector<u64> out(bit_sizes.size(), 0);
    std::vector<char> done(bit_sizes.size(), 0);
    const u64 step = 2 * n;
    for (size_t i = 0; i < bit_sizes.size(); ++i) {
        if (done[i]) continue;
        int bits = bit_sizes[i];
        if (bits < 2 || bits > 61) throw std::invalid_argument("bit size out of range");
        std::vector<size_t> slots;
        for (size_t j = i; j < bit_sizes.size(); ++j)
            if (bit_sizes[j] == bits) slots.push_back(j), done[j] = 1;
        std::vector<u64> found;
        u64 cand = ((u64) 1 << bits);
        if (cand < step) throw std::logic_error("failed to find enough qualifying primes");
        cand = cand - step + 1;
        const u64 floor_v = (u64) 1 << (bits - 1);
        while (found.size() < slots.size() && cand > floor_v) {
            if (is_prime(cand)) found.push_back(cand);
            cand -= step;
        }
        if (found.size() < slots.size()) throw std::logic_error("failed to find enough qualifying primes");
        for (size_t k = 0; k < slots.size(); ++k) out[slots[k]] = found[slots.size() - 1 - k];
    }
    return out;
}

// smallest primitive `degree`-th root of unity mod q (degree a power of two dividing q-1)
inline u64 minimal_primitive_root(u64 degree, u64 q) {
    if ((q - 1) % degree != 0) throw std::invalid_argument("q - 1 not divisible by 2N");
    const u64 cofactor = (q - 1) / degree;
    u64 g = 0;
    for (u64 x = 2; x < q && !g; ++x) {
        u64 cand = powmod(x, cofactor, q);
        if (powmod(cand, degree / 2, q) == q - 1) g = cand;
    }
    if (!g) throw std::invalid_argument("no primitive root");
    // all primitive roots are the odd powers of g
    const u64 g2 = mulmod(g, g, q);
    u64 best = g, cur = g;
    for (u64 i = 1; i < degree / 2; ++i) {
        cur = mulmod(cur, g2, q);
        if (cur < best) best = cur;
    }
    return best;
}

inline uint32_t bit_reverse(uint32_t x, int bits) {
    uint32_t r = 0;
    for (int i = 0; i < bits; ++i) r = (r << 1) | ((x >> i) & 1);
    return r;
}

// prod_{k != skip} base[k] mod p   (skip < 0: full product)
inline u64 product_mod(const std::vector<u64> &base, int skip, u64 p) {
    u64 r = 1 % p;
    for (int k = 0; k < (int) base.size(); ++k)
        if (k != skip) r = mulmod(r, base[k] % p, p);
    return r;
}

// NTT-friendly primes below `upper` (itself 1 mod 2N), descending (get_primes_below, numth.cu:235-263)
inline std::vector<u64> primes_below(u64 n, u64 upper, size_t count) {
    const u64 step = 2 * n;
    if (upper < step) throw std::logic_error("failed to find enough qualifying primes 1");
    const int bits = 64 - __builtin_clzll(upper);
    const u64 floor_v = (u64) 1 << (bits - 1);
    std::vector<u64> out;
    for (u64 v = upper - step; out.size() < count && v > floor_v; v -= step)
        if (is_prime(v)) out.push_back(v);
    if (out.size() < count) throw std::logic_error("failed to find enough qualifying primes 2");
    return out;
}

// little-endian multi-word integers for the HPS scale-and-round tables (t * R * x / s needs ~(size_R + 2) words)
struct BigUint {
    std::vector<u64> w{1};
    void mul_word(u64 m) {
        u64 carry = 0;
        for (auto &x : w) {
            const u128 t = (u128) x * m + carry;
            x = (u64) t;
            carry = (u64) (t >> 64);
        }
        if (carry) w.push_back(carry);
    }
    u64 divmod_word(u64 d) {   // *this <- floor(*this / d), returns the remainder
        u64 rem = 0;
        for (size_t k = w.size(); k-- > 0;) {
            const u128 cur = ((u128) rem << 64) | w[k];
            w[k] = (u64) (cur / d);
            rem = (u64) (cur % d);
        }
        while (w.size() > 1 && w.back() == 0) w.pop_back();
        return rem;
    }
    u64 mod_word(u64 d) const {
        u64 rem = 0;
        for (size_t k = w.size(); k-- > 0;) rem = (u64) ((((u128) rem << 64) | w[k]) % d);
        return rem;
    }
};

// bit length of the product of `base` (base_Q.big_modulus significant bits, rns.cu:400-406)
inline int product_bits(const std::vector<u64> &base) {
    std::vector<u64> acc{1};
    for (u64 p : base) {
        u64 carry = 0;
        for (auto &w : acc) {
            const u128 t = (u128) w * p + carry;
            w = (u64) t;
            carry = (u64) (t >> 64);
        }
        if (carry) acc.push_back(carry);
    }
    return (int) (acc.size() - 1) * 64 + (64 - __builtin_clzll(acc.back()));
}

} // namespace pfhe::host

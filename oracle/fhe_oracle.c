/*
 * fhe_oracle.c -- CPU restatement of the Phantom-FHE RNS hot path.  TEST INFRASTRUCTURE ONLY
 * (see fhe_oracle.h for who may load this and how its parity is pinned).
 *
 * Plain C99 + unsigned __int128.  Everything is written as "obviously correct" modular arithmetic that
 * produces canonical residues; the NTT uses SEAL-style Harvey lazy butterflies over the same bit-reversed
 * tables the reference's host code generates, because that is also the "SEAL-style CPU path" timed as
 * cpu_baseline by bench.py (OpenMP over limbs).
 */
#include "fhe_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef uint64_t u64;

static int g_threads = 1;
void orc_set_threads(int nthreads) { g_threads = nthreads < 1 ? 1 : nthreads; }
int orc_get_threads(void) { return g_threads; }
/* diagnosis only (tests/test_bfv_hps_alpha_case.py): 1 = reduce the UNREDUCED alpha under every r_j in the HPS scale-and-round
 * step instead of restating the reference's limb-after-limb re-reduction (rns.cu:1699-1733) */
static int g_hps_alpha_per_limb = 0;
void orc_set_hps_alpha_per_limb(int on) { g_hps_alpha_per_limb = on != 0; }

/* ------------------------------------------------------------------------------------------------------
 * scalar modular arithmetic
 * ---------------------------------------------------------------------------------------------------- */
u64 orc_mulmod(u64 a, u64 b, u64 q) { return (u64)(((u128)a * b) % q); }

static u64 powmod(u64 a, u64 e, u64 q) {
    u64 r = 1 % q;
    a %= q;
    while (e) {
        if (e & 1) r = orc_mulmod(r, a, q);
        a = orc_mulmod(a, a, q);
        e >>= 1;
    }
    return r;
}

/* try_invert_uint_mod (src/host/uintarithsmallmod.cu): a^-1 mod q, q need not be prime -> extended Euclid */
u64 orc_invmod(u64 a, u64 q) {
    __int128 t0 = 0, t1 = 1;
    u64 r0 = q, r1 = a % q;
    while (r1) {
        u64 k = r0 / r1;
        __int128 t2 = t0 - (__int128)k * t1;
        t0 = t1;
        t1 = t2;
        u64 r2 = r0 - k * r1;
        r0 = r1;
        r1 = r2;
    }
    if (r0 != 1) return 0; /* not invertible */
    if (t0 < 0) t0 += q;
    return (u64)t0;
}

/* compute_shoup: floor(w * 2^64 / q), include/host/uintarithsmallmod.h:119-124 */
u64 orc_shoup(u64 w, u64 q) { return (u64)((((u128)w) << 64) / q); }

/* Modulus::set_value, src/host/modulus.cu:28-41: floor(2^128 / q) as two words + remainder */
void orc_barrett_ratio(u64 q, u64 ratio[3]) {
    /* long division of [0,0,1] (base 2^64) by q */
    u64 rem = 1; /* top word 1 < q */
    u128 cur = ((u128)rem << 64);
    u64 q1 = (u64)(cur / q);
    rem = (u64)(cur % q);
    cur = ((u128)rem << 64);
    u64 q0 = (u64)(cur / q);
    rem = (u64)(cur % q);
    ratio[0] = q0;
    ratio[1] = q1;
    ratio[2] = rem;
}

/* Shoup multiplication, canonical result (uintmodmath.cuh:207-216) */
static inline u64 mul_shoup(u64 x, u64 w, u64 ws, u64 q) {
    u64 hi = (u64)(((u128)x * ws) >> 64);
    u64 r = x * w - hi * q;
    return r >= q ? r - q : r;
}
/* lazy variant -> [0, 2q) (uintmodmath.cuh:226-231) */
static inline u64 mul_shoup_lazy(u64 x, u64 w, u64 ws, u64 q) {
    u64 hi = (u64)(((u128)x * ws) >> 64);
    return x * w - hi * q;
}

static inline u64 addmod(u64 a, u64 b, u64 q) {
    u64 r = a + b;
    return r >= q ? r - q : r;
}
static inline u64 submod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }

/* is_prime: the reference runs Miller-Rabin with random bases (src/host/numth.cu:160-204); the outcome for
 * 64-bit inputs is the deterministic primality predicate, computed here with the first 12 prime bases. */
int orc_is_prime(u64 v) {
    static const u64 bases[12] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    if (v < 2) return 0;
    for (int i = 0; i < 12; i++) {
        if (v == bases[i]) return 1;
        if (v % bases[i] == 0) return 0;
    }
    u64 d = v - 1;
    int r = 0;
    while (!(d & 1)) {
        d >>= 1;
        r++;
    }
    for (int i = 0; i < 12; i++) {
        u64 x = powmod(bases[i], d, v);
        if (x == 1 || x == v - 1) continue;
        int comp = 1;
        for (int k = 1; k < r; k++) {
            x = orc_mulmod(x, x, v);
            if (x == v - 1) {
                comp = 0;
                break;
            }
        }
        if (comp) return 0;
    }
    return 1;
}

/* CoeffModulus::Create (src/host/modulus.cu:79-110) over get_primes (src/host/numth.cu:207-233):
 * per distinct bit size, primes are found downwards from 2^bits - 2n + 1 in steps of 2n; the list is then
 * consumed from the BACK, so for repeated sizes the later list position gets the larger prime. */
int orc_create_primes(u64 n, const int *bit_sizes, int count, u64 *out) {
    if (count <= 0 || count > 64) return -1;
    int done[64];
    memset(done, 0, sizeof(done));
    for (int i = 0; i < count; i++) {
        if (done[i]) continue;
        int bits = bit_sizes[i];
        if (bits < 2 || bits > 61) return -1;
        int pos[64], npos = 0;
        for (int j = i; j < count; j++)
            if (bit_sizes[j] == bits) {
                pos[npos++] = j;
                done[j] = 1;
            }
        u64 found[64];
        int nf = 0;
        u64 factor = 2 * n;
        u64 value = ((u64)1 << bits);
        if (value < factor) return -1;
        value = value - factor + 1;
        u64 lower = (u64)1 << (bits - 1);
        while (nf < npos && value > lower) {
            if (orc_is_prime(value)) found[nf++] = value;
            value -= factor;
        }
        if (nf < npos) return -1;
        /* result.emplace_back(prime_table[size].back()); pop_back() */
        for (int k = 0; k < npos; k++) out[pos[k]] = found[npos - 1 - k];
    }
    return 0;
}

/* try_minimal_primitive_root (src/host/numth.cu:309-331) */
u64 orc_minimal_primitive_root(u64 degree, u64 q) {
    if ((q - 1) % degree) return 0;
    u64 quot = (q - 1) / degree;
    u64 root = 0;
    for (u64 x = 2; x < q; x++) {
        u64 g = powmod(x, quot, q);
        if (powmod(g, degree >> 1, q) == q - 1) {
            root = g;
            break;
        }
    }
    if (!root) return 0;
    u64 gsq = orc_mulmod(root, root, q);
    u64 cur = root, best = root;
    for (u64 i = 0; i < degree / 2; i++) {
        if (cur < best) best = cur;
        cur = orc_mulmod(cur, gsq, q);
    }
    return best;
}

static inline uint32_t bitrev32(uint32_t x, int bits) {
    uint32_t r = 0;
    for (int i = 0; i < bits; i++) {
        r = (r << 1) | (x & 1);
        x >>= 1;
    }
    return r;
}

/* ------------------------------------------------------------------------------------------------------
 * context
 * ---------------------------------------------------------------------------------------------------- */
struct orc_ctx {
    int scheme;
    u64 n;
    int logn;
    int size_QP, size_P, size_Q;
    u64 t;
    u64 *primes;                 /* [size_QP] */
    u64 *tw, *tws, *itw, *itws;  /* [size_QP][n] */
    u64 *ninv, *ninvs;           /* [size_QP] */
};

orc_ctx *orc_create(int scheme, u64 n, const u64 *primes, int size_QP, int size_P, u64 t) {
    int logn = 0;
    while (((u64)1 << logn) < n) logn++;
    if (((u64)1 << logn) != n || size_QP < 1 || size_P < 0 || size_P >= size_QP + (size_P == 0)) return NULL;
    orc_ctx *c = (orc_ctx *)calloc(1, sizeof(orc_ctx));
    c->scheme = scheme;
    c->n = n;
    c->logn = logn;
    c->size_QP = size_QP;
    c->size_P = size_P;
    c->size_Q = size_QP - size_P;
    c->t = t;
    c->primes = (u64 *)malloc(sizeof(u64) * size_QP);
    memcpy(c->primes, primes, sizeof(u64) * size_QP);
    size_t tab = (size_t)size_QP * n;
    c->tw = (u64 *)malloc(tab * 8);
    c->tws = (u64 *)malloc(tab * 8);
    c->itw = (u64 *)malloc(tab * 8);
    c->itws = (u64 *)malloc(tab * 8);
    c->ninv = (u64 *)malloc(sizeof(u64) * size_QP);
    c->ninvs = (u64 *)malloc(sizeof(u64) * size_QP);
    int ok = 1;
#pragma omp parallel for num_threads(g_threads) schedule(dynamic)
    for (int i = 0; i < size_QP; i++) {
        u64 q = primes[i];
        /* NTT::NTT, src/host/ntt.cu:10-56 */
        u64 root = orc_minimal_primitive_root(2 * n, q);
        if (!root) {
            ok = 0;
            continue;
        }
        u64 iroot = orc_invmod(root, q);
        u64 *tw = c->tw + (size_t)i * n, *tws = c->tws + (size_t)i * n;
        u64 *itw = c->itw + (size_t)i * n, *itws = c->itws + (size_t)i * n;
        u64 p = root, ip = iroot;
        for (u64 k = 1; k < n; k++) {
            uint32_t r = bitrev32((uint32_t)k, logn);
            tw[r] = p;
            tws[r] = orc_shoup(p, q);
            itw[r] = ip;
            itws[r] = orc_shoup(ip, q);
            p = orc_mulmod(p, root, q);
            ip = orc_mulmod(ip, iroot, q);
        }
        tw[0] = 1;
        tws[0] = orc_shoup(1, q);
        itw[0] = 1;
        itws[0] = orc_shoup(1, q);
        u64 ninv = orc_invmod(n % q, q);
        c->ninv[i] = ninv;
        c->ninvs[i] = orc_shoup(ninv, q);
        /* inv_root_powers_[1] *= n^-1 (src/host/ntt.cu:53-55) */
        itw[1] = orc_mulmod(itw[1], ninv, q);
        itws[1] = orc_shoup(itw[1], q);
    }
    if (!ok) {
        orc_destroy(c);
        return NULL;
    }
    return c;
}

void orc_destroy(orc_ctx *c) {
    if (!c) return;
    free(c->primes);
    free(c->tw);
    free(c->tws);
    free(c->itw);
    free(c->itws);
    free(c->ninv);
    free(c->ninvs);
    free(c);
}

u64 orc_n(const orc_ctx *c) { return c->n; }
int orc_size_QP(const orc_ctx *c) { return c->size_QP; }
int orc_size_P(const orc_ctx *c) { return c->size_P; }
const u64 *orc_twiddle(const orc_ctx *c, int i) { return c->tw + (size_t)i * c->n; }
const u64 *orc_twiddle_shoup(const orc_ctx *c, int i) { return c->tws + (size_t)i * c->n; }
const u64 *orc_itwiddle(const orc_ctx *c, int i) { return c->itw + (size_t)i * c->n; }
const u64 *orc_itwiddle_shoup(const orc_ctx *c, int i) { return c->itws + (size_t)i * c->n; }
u64 orc_n_inv(const orc_ctx *c, int i) { return c->ninv[i]; }
u64 orc_prime(const orc_ctx *c, int i) { return c->primes[i]; }

/* std::mt19937_64 (ISO C++ [rand.predef]: w=64, n=312, m=156, r=31, a=0xb5026f5aa96619e9, u=29,
 * d=0x5555555555555555, s=17, b=0x71d67fffeda60000, t=37, c=0xfff7eee000000000, l=43, f=6364136223846793005) */
void orc_mt19937_64_fill(u64 seed, u64 q, u64 *out, size_t count, int mode) {
    u64 mt[312];
    int idx = 312;
    mt[0] = seed;
    for (int i = 1; i < 312; i++) mt[i] = 6364136223846793005ULL * (mt[i - 1] ^ (mt[i - 1] >> 62)) + (u64)i;
    int bits = 0;
    while (bits < 64 && (q >> bits)) bits++;
    u64 mask = bits >= 64 ? ~(u64)0 : (((u64)1 << bits) - 1);
    size_t produced = 0;
    while (produced < count) {
        if (idx >= 312) {
            for (int i = 0; i < 312; i++) {
                u64 x = (mt[i] & 0xFFFFFFFF80000000ULL) | (mt[(i + 1) % 312] & 0x7FFFFFFFULL);
                u64 xa = x >> 1;
                if (x & 1) xa ^= 0xB5026F5AA96619E9ULL;
                mt[i] = mt[(i + 156) % 312] ^ xa;
            }
            idx = 0;
        }
        u64 y = mt[idx++];
        y ^= (y >> 29) & 0x5555555555555555ULL;
        y ^= (y << 17) & 0x71D67FFFEDA60000ULL;
        y ^= (y << 37) & 0xFFF7EEE000000000ULL;
        y ^= (y >> 43);
        if (mode == 0) {
            out[produced++] = y % q;
        } else {
            y &= mask;
            if (y < q) out[produced++] = y;
        }
    }
}

/* key-level prime index of position j in the packed data-level base Ql ∪ P (fntt_2d.cu:434-436) */
static inline int qlp_index(const orc_ctx *c, int l, int j) { return j < l ? j : c->size_Q + (j - l); }

/* ------------------------------------------------------------------------------------------------------
 * NTT
 * ---------------------------------------------------------------------------------------------------- */
/* forward negacyclic NTT, natural -> bit-reversed order (ntt_1d.cu:41-73; butterfly.cuh:10-22 semantics:
 * Harvey lazy CT butterfly, x,y in [0,4q) -> [0,4q)), final reduction to [0,q) (fntt_2d.cu:187-193). */
static void ntt_fwd_limb(u64 *a, u64 n, const u64 *tw, const u64 *tws, u64 q) {
    u64 two_q = 2 * q;
    for (u64 m = 1; m < n; m <<= 1) {
        u64 gap = n / (2 * m);
        for (u64 i = 0; i < m; i++) {
            u64 w = tw[m + i], ws = tws[m + i];
            u64 *x = a + 2 * i * gap, *y = x + gap;
            for (u64 j = 0; j < gap; j++) {
                u64 X = x[j];
                if (X >= two_q) X -= two_q;
                u64 T = mul_shoup_lazy(y[j], w, ws, q);
                x[j] = X + T;
                y[j] = X + two_q - T;
            }
        }
    }
    for (u64 j = 0; j < n; j++) {
        u64 v = a[j];
        if (v >= two_q) v -= two_q;
        if (v >= q) v -= q;
        a[j] = v;
    }
}

/* inverse negacyclic NTT, bit-reversed -> natural (ntt_1d.cu:220-254; butterfly.cuh:28-37: GS butterfly,
 * x,y in [0,2q) -> [0,2q)); lower half times n^-1, the upper half got it through itw[1]
 * (intt_2d.cu:195-198); final reduction to [0,q) (intt_2d.cu:203). */
static void ntt_inv_limb(u64 *a, u64 n, const u64 *itw, const u64 *itws, u64 ninv, u64 ninvs, u64 q) {
    u64 two_q = 2 * q;
    for (u64 m = n >> 1; m >= 1; m >>= 1) {
        u64 gap = n / (2 * m);
        for (u64 i = 0; i < m; i++) {
            u64 w = itw[m + i], ws = itws[m + i];
            u64 *x = a + 2 * i * gap, *y = x + gap;
            for (u64 j = 0; j < gap; j++) {
                u64 X = x[j], Y = y[j];
                u64 S = X + Y;
                if (S >= two_q) S -= two_q;
                u64 D = X + two_q - Y;
                x[j] = S;
                y[j] = mul_shoup_lazy(D, w, ws, q);
            }
        }
    }
    for (u64 j = 0; j < n / 2; j++) a[j] = mul_shoup(a[j], ninv, ninvs, q);
    for (u64 j = n / 2; j < n; j++) {
        u64 v = a[j];
        if (v >= q) v -= q;
        a[j] = v;
    }
}

void orc_ntt_forward(const orc_ctx *c, u64 *data, int limbs, const int *table_idx) {
#pragma omp parallel for num_threads(g_threads) schedule(dynamic)
    for (int i = 0; i < limbs; i++) {
        int k = table_idx[i];
        ntt_fwd_limb(data + (size_t)i * c->n, c->n, c->tw + (size_t)k * c->n, c->tws + (size_t)k * c->n,
                     c->primes[k]);
    }
}

void orc_ntt_inverse(const orc_ctx *c, u64 *data, int limbs, const int *table_idx) {
#pragma omp parallel for num_threads(g_threads) schedule(dynamic)
    for (int i = 0; i < limbs; i++) {
        int k = table_idx[i];
        ntt_inv_limb(data + (size_t)i * c->n, c->n, c->itw + (size_t)k * c->n, c->itws + (size_t)k * c->n,
                     c->ninv[k], c->ninvs[k], c->primes[k]);
    }
}

/* fnwt_1d[_opt] / inwt_1d[_opt] (reference src/ntt/ntt_1d.cu:146-292): single-block transforms for dim <= 2048 on
 * CALLER-SUPPLIED tables.  Limb i of the call is absolute index start + i in inout, in the tables, in the modulus and
 * scalar arrays (ntt_1d.cu:27-28,208-209).  Inverse: the lower half is multiplied by scalar[idx], the upper half only
 * carries what the caller folded into itwiddles[1] (ntt_1d.cu:245-248). */
void orc_fnwt_1d(u64 *inout, const u64 *tw, const u64 *tws, const u64 *q, u64 dim, int count, int start) {
    for (int i = start; i < start + count; i++)
        ntt_fwd_limb(inout + (size_t)i * dim, dim, tw + (size_t)i * dim, tws + (size_t)i * dim, q[i]);
}
void orc_inwt_1d(u64 *inout, const u64 *itw, const u64 *itws, const u64 *q, const u64 *scalar, const u64 *scalar_shoup,
                 u64 dim, int count, int start) {
    for (int i = start; i < start + count; i++)
        ntt_inv_limb(inout + (size_t)i * dim, dim, itw + (size_t)i * dim, itws + (size_t)i * dim, scalar[i],
                     scalar_shoup[i], q[i]);
}

/* ------------------------------------------------------------------------------------------------------
 * dyadic kernels
 * ---------------------------------------------------------------------------------------------------- */
void orc_tensor_2x2(const orc_ctx *c, const u64 *a, const u64 *b, u64 *out, int l) {
    size_t n = c->n, poly = (size_t)l * n;
#pragma omp parallel for num_threads(g_threads)
    for (int i = 0; i < l; i++) {
        u64 q = c->primes[i];
        for (size_t j = 0; j < n; j++) {
            size_t k = (size_t)i * n + j;
            u64 a0 = a[k], a1 = a[k + poly], b0 = b[k], b1 = b[k + poly];
            u64 d0 = orc_mulmod(a0, b0, q);
            u64 d2 = orc_mulmod(a1, b1, q);
            /* (c0+c1)(c0'+c1') - d0 - d2 (polymath.cu:489-493); the sums are taken unreduced there too */
            u64 d1 = (u64)(((u128)(a0 + a1) * (b0 + b1)) % q);
            d1 = submod(submod(d1, d0, q), d2, q);
            out[k] = d0;
            out[k + poly] = d1;
            out[k + 2 * poly] = d2;
        }
    }
}

void orc_tensor_square_2x2(const orc_ctx *c, const u64 *a, u64 *out, int l) {
    size_t n = c->n, poly = (size_t)l * n;
#pragma omp parallel for num_threads(g_threads)
    for (int i = 0; i < l; i++) {
        u64 q = c->primes[i];
        for (size_t j = 0; j < n; j++) {
            size_t k = (size_t)i * n + j;
            u64 a0 = a[k], a1 = a[k + poly];
            u64 d0 = orc_mulmod(a0, a0, q);
            u64 d1 = (u64)((((u128)a0 * a1) << 1) % q);
            u64 d2 = orc_mulmod(a1, a1, q);
            out[k] = d0;
            out[k + poly] = d1;
            out[k + 2 * poly] = d2;
        }
    }
}

/* tensor_prod_mxn_rns_poly (reference src/polymath.cu:546-594): out[j] = sum_{i1+i2=j} a[i1] * b[i2] mod q, j <
 * sa + sb - 1, every sum accumulated in 128 bits and reduced once.  out may alias a (the reference's in-place form):
 * all operands of a coefficient are read before its results are written. */
void orc_tensor_mxn(const orc_ctx *c, const u64 *a, int sa, const u64 *b, int sb, u64 *out, int l) {
    size_t n = c->n, poly = (size_t)l * n;
    int so = sa + sb - 1;
#pragma omp parallel for num_threads(g_threads)
    for (int i = 0; i < l; i++) {
        u64 q = c->primes[i];
        u64 c1[ORC_MAX_CT], c2[ORC_MAX_CT], r[2 * ORC_MAX_CT];
        for (size_t x = 0; x < n; x++) {
            size_t k = (size_t)i * n + x;
            for (int u = 0; u < sa; u++) c1[u] = a[k + u * poly];
            for (int u = 0; u < sb; u++) c2[u] = b[k + u * poly];
            for (int j = 0; j < so; j++) {
                int last1 = j < sa - 1 ? j : sa - 1, first2 = j < sb - 1 ? j : sb - 1, first1 = j - first2;
                u128 acc = 0;
                for (int u = 0; u <= last1 - first1; u++) acc += (u128)c1[first1 + u] * c2[first2 - u];
                r[j] = (u64)(acc % q);
            }
            for (int j = 0; j < so; j++) out[k + j * poly] = r[j];
        }
    }
}

#define ELEMENTWISE(name, expr)                                                                          \
    void name(const orc_ctx *c, const u64 *a, const u64 *b, u64 *out, int l) {                           \
        size_t n = c->n;                                                                                 \
        for (int i = 0; i < l; i++) {                                                                    \
            u64 q = c->primes[i];                                                                        \
            for (size_t j = 0; j < n; j++) {                                                             \
                size_t k = (size_t)i * n + j;                                                            \
                out[k] = (expr);                                                                         \
            }                                                                                            \
        }                                                                                                \
    }
ELEMENTWISE(orc_poly_add, addmod(a[k], b[k], q))
ELEMENTWISE(orc_poly_sub, submod(a[k], b[k], q))
ELEMENTWISE(orc_poly_mul, orc_mulmod(a[k], b[k], q))

void orc_poly_negate(const orc_ctx *c, const u64 *a, u64 *out, int l) {
    size_t n = c->n;
    for (int i = 0; i < l; i++) {
        u64 q = c->primes[i];
        for (size_t j = 0; j < n; j++) {
            size_t k = (size_t)i * n + j;
            out[k] = a[k] ? q - a[k] : 0;
        }
    }
}

/* ------------------------------------------------------------------------------------------------------
 * fast base conversion helpers (src/host/rns.cu:282-337,438-497; src/rns_bconv.cu:22-60,143-168)
 * ---------------------------------------------------------------------------------------------------- */
/* punctured product of ibase except i, modulo p */
static u64 qhat_mod(const u64 *ibase, int ni, int i, u64 p) {
    u64 r = 1 % p;
    for (int k = 0; k < ni; k++)
        if (k != i) r = orc_mulmod(r, ibase[k] % p, p);
    return r;
}

/* out[j][x] = sum_i (in[i][x] * qhatinv_i mod q_i) * (qhat_i mod p_j) mod p_j  (bConv_BEHZ,
 * rns_bconv.cu:212-229).  If prescaled != 0 the input has already been multiplied by qhatinv_i. */
static void bconv(const u64 *ibase, int ni, const u64 *obase, int no, const u64 *const *in, u64 *const *out,
                  size_t n, int prescaled) {
    u64 hinv[64], hinvs[64];
    for (int i = 0; i < ni; i++) {
        hinv[i] = orc_invmod(qhat_mod(ibase, ni, i, ibase[i]), ibase[i]);
        hinvs[i] = orc_shoup(hinv[i], ibase[i]);
    }
#pragma omp parallel for num_threads(g_threads)
    for (int j = 0; j < no; j++) {
        u64 p = obase[j];
        u64 mat[64];
        for (int i = 0; i < ni; i++) mat[i] = qhat_mod(ibase, ni, i, p);
        for (size_t x = 0; x < n; x++) {
            u128 acc = 0; /* <= 64 terms of < 2^122 each is not representable in general, reduce per term */
            for (int i = 0; i < ni; i++) {
                u64 y = prescaled ? in[i][x] : mul_shoup(in[i][x], hinv[i], hinvs[i], ibase[i]);
                acc += (u128)y * mat[i];
                if ((i & 3) == 3) acc %= p;
            }
            out[j][x] = (u64)(acc % p);
        }
    }
}

int orc_beta(const orc_ctx *c, int l) { return c->size_P ? (l + c->size_P - 1) / c->size_P : 0; }

/* ------------------------------------------------------------------------------------------------------
 * key switching
 * ---------------------------------------------------------------------------------------------------- */
void orc_modup(const orc_ctx *c, int l, const u64 *cks, u64 *t_mod_up) {
    size_t n = c->n;
    int alpha = c->size_P, m = l + alpha, beta = orc_beta(c, l);
    int is_bfv = c->scheme == ORC_SCHEME_BFV;

    /* t_cks: coefficient-domain copy of cks (rns_bconv.cu:544-560) */
    u64 *t_cks = (u64 *)malloc((size_t)l * n * 8);
    memcpy(t_cks, cks, (size_t)l * n * 8);
    if (!is_bfv) {
        int idx[64];
        for (int i = 0; i < l; i++) idx[i] = i;
        orc_ntt_inverse(c, t_cks, l, idx);
    }

    for (int d = 0; d < beta; d++) {
        int start = alpha * d;
        int size_part = (d == beta - 1) ? (l - alpha * (beta - 1)) : alpha;
        int end = start + size_part;
        u64 *dst = t_mod_up + (size_t)d * m * n;

        u64 ibase[64], obase[64];
        const u64 *in[64];
        u64 *out[64];
        int oidx[64];
        for (int i = 0; i < size_part; i++) {
            ibase[i] = c->primes[start + i];
            in[i] = t_cks + (size_t)(start + i) * n;
        }
        int no = 0;
        for (int j = 0; j < m; j++) {
            if (j >= start && j < end) continue;
            int k = qlp_index(c, l, j);
            obase[no] = c->primes[k];
            out[no] = dst + (size_t)j * n;
            oidx[no] = k;
            no++;
        }
        /* alpha == 1: copy / reduce (modup_bconv_single_p_kernel, rns_bconv.cu:432-453) == bconv with
         * qhat = qhatinv = 1; alpha > 1: scale by partQlHatInv (rns_bconv.cu:558 / :598) then matmul
         * (bconv_matmul_padded_unroll2_kernel, :455-485) */
        bconv(ibase, size_part, obase, no, in, out, n, 0);

        /* the digit's own limbs: raw copy of cks (modup_copy_partQl_kernel :522-528 / single_p :450) */
        for (int j = start; j < end; j++) memcpy(dst + (size_t)j * n, cks + (size_t)j * n, n * 8);

        if (!is_bfv) {
            /* NTT on the converted limbs only (..._exclude_range, rns_bconv.cu:618) */
            for (int k = 0; k < no; k++) orc_ntt_forward(c, out[k], 1, &oidx[k]);
        } else {
            /* BFV: all l+alpha limbs go to NTT form (rns_bconv.cu:622) */
            for (int j = 0; j < m; j++) {
                int k = qlp_index(c, l, j);
                orc_ntt_forward(c, dst + (size_t)j * n, 1, &k);
            }
        }
    }
    free(t_cks);
}

void orc_inner_prod(const orc_ctx *c, int l, const u64 *t_mod_up, const u64 *evk, u64 *cx) {
    size_t n = c->n;
    int alpha = c->size_P, m = l + alpha, beta = orc_beta(c, l);
    size_t qp_n = (size_t)c->size_QP * n, m_n = (size_t)m * n;
#pragma omp parallel for num_threads(g_threads)
    for (int j = 0; j < m; j++) {
        int twr = qlp_index(c, l, j);
        u64 q = c->primes[twr];
        for (size_t x = 0; x < n; x++) {
            u128 a0 = 0, a1 = 0;
            for (int d = 0; d < beta; d++) {
                u64 v = t_mod_up[(size_t)d * m_n + (size_t)j * n + x];
                const u64 *k = evk + (size_t)d * 2 * qp_n + (size_t)twr * n + x;
                a0 = (a0 + (u128)v * k[0]) % q;
                a1 = (a1 + (u128)v * k[qp_n]) % q;
            }
            cx[(size_t)j * n + x] = (u64)a0;
            cx[m_n + (size_t)j * n + x] = (u64)a1;
        }
    }
}

/* P -> Ql conversion used by mod-down: delta[j] = sum_i (x_i * Phatinv_i) * (Phat_i mod q_j)
 * (alpha == 1: x mod q_j, moddown_bconv_single_p_kernel rns_bconv.cu:691-707) */
static void p_to_ql(const orc_ctx *c, int l, const u64 *xp, u64 *delta) {
    size_t n = c->n;
    u64 ibase[64], obase[64];
    const u64 *in[64];
    u64 *out[64];
    for (int i = 0; i < c->size_P; i++) {
        ibase[i] = c->primes[c->size_Q + i];
        in[i] = xp + (size_t)i * n;
    }
    for (int j = 0; j < l; j++) {
        obase[j] = c->primes[j];
        out[j] = delta + (size_t)j * n;
    }
    bconv(ibase, c->size_P, obase, l, in, out, n, 0);
}

static u64 bigP_mod(const orc_ctx *c, u64 q) {
    u64 r = 1 % q;
    for (int i = 0; i < c->size_P; i++) r = orc_mulmod(r, c->primes[c->size_Q + i] % q, q);
    return r;
}

void orc_moddown_from_ntt(const orc_ctx *c, int l, u64 *cx_i, u64 *ct_i) {
    size_t n = c->n;
    int alpha = c->size_P, m = l + alpha;
    int idx[64];
    u64 *delta = (u64 *)malloc((size_t)l * n * 8);

    if (c->scheme == ORC_SCHEME_CKKS) {
        for (int i = 0; i < alpha; i++) idx[i] = c->size_Q + i;
        orc_ntt_inverse(c, cx_i + (size_t)l * n, alpha, idx); /* rns_bconv.cu:788 */
    } else {
        for (int j = 0; j < m; j++) idx[j] = qlp_index(c, l, j);
        orc_ntt_inverse(c, cx_i, m, idx); /* :792 */
    }
    p_to_ql(c, l, cx_i + (size_t)l * n, delta); /* :796-802 */

    if (c->scheme == ORC_SCHEME_BGV) {
        /* base_P_to_t_conv_.bConv_BEHZ + bgv_moddown_kernel (rns_bconv.cu:636-652,804-817) */
        u64 t = c->t;
        u64 *cp_t = (u64 *)malloc(n * 8);
        u64 ibase[64];
        const u64 *in[64];
        for (int i = 0; i < alpha; i++) {
            ibase[i] = c->primes[c->size_Q + i];
            in[i] = cx_i + (size_t)(l + i) * n;
        }
        u64 *outp[1] = {cp_t};
        bconv(ibase, alpha, &t, 1, in, outp, n, 0);
        u64 pinv_t = orc_invmod(bigP_mod(c, t), t);
        for (int j = 0; j < l; j++) {
            u64 q = c->primes[j];
            u64 P = bigP_mod(c, q), Pinv = orc_invmod(P, q);
            for (size_t x = 0; x < n; x++) {
                u64 tmp = orc_mulmod(cp_t[x], pinv_t, t);
                u64 corr = orc_mulmod(tmp, P, q);
                u64 v = submod(cx_i[(size_t)j * n + x], delta[(size_t)j * n + x], q);
                v = addmod(v, corr, q);
                ct_i[(size_t)j * n + x] = orc_mulmod(v, Pinv, q);
            }
        }
        for (int j = 0; j < l; j++) idx[j] = j;
        orc_ntt_forward(c, ct_i, l, idx);
        free(cp_t);
    } else {
        if (c->scheme == ORC_SCHEME_CKKS) {
            for (int j = 0; j < l; j++) idx[j] = j;
            orc_ntt_forward(c, delta, l, idx); /* fused in nwt_2d_radix8_forward_inplace_fuse_moddown :820 */
        }
        /* (cx - delta) * P^-1 mod q_j  (ntt_moddown.cu:210-213 / moddown_kernel rns_bconv.cu:680-689) */
        for (int j = 0; j < l; j++) {
            u64 q = c->primes[j];
            u64 Pinv = orc_invmod(bigP_mod(c, q), q);
            for (size_t x = 0; x < n; x++) {
                u64 v = submod(cx_i[(size_t)j * n + x], delta[(size_t)j * n + x], q);
                ct_i[(size_t)j * n + x] = orc_mulmod(v, Pinv, q);
            }
        }
    }
    free(delta);
}

void orc_keyswitch(const orc_ctx *c, int l, u64 *ct, const u64 *c2, const u64 *evk) {
    size_t n = c->n;
    int m = l + c->size_P, beta = orc_beta(c, l);
    u64 *t_mod_up = (u64 *)malloc((size_t)beta * m * n * 8);
    u64 *cx = (u64 *)malloc((size_t)2 * m * n * 8);
    u64 *res = (u64 *)malloc((size_t)l * n * 8);
    orc_modup(c, l, c2, t_mod_up);
    orc_inner_prod(c, l, t_mod_up, evk, cx);
    for (int k = 0; k < 2; k++) {
        /* the reference calls moddown_from_NTT(cx_i, cx_i, ...): the result lands in cx_i[0..l) and is then
         * added to ct_i (add_to_ct_kernel, rns_bconv.cu:763-769) */
        orc_moddown_from_ntt(c, l, cx + (size_t)k * m * n, res);
        orc_poly_add(c, ct + (size_t)k * l * n, res, ct + (size_t)k * l * n, l);
    }
    free(t_mod_up);
    free(cx);
    free(res);
}

void orc_multiply_relin(const orc_ctx *c, int l, const u64 *ct1, const u64 *ct2, const u64 *rlk, u64 *out) {
    size_t poly = (size_t)l * c->n;
    u64 *d = (u64 *)malloc(3 * poly * 8);
    orc_tensor_2x2(c, ct1, ct2, d, l);
    orc_keyswitch(c, l, d, d + 2 * poly, rlk);
    memcpy(out, d, 2 * poly * 8);
    free(d);
}

/* ------------------------------------------------------------------------------------------------------
 * Galois
 * ---------------------------------------------------------------------------------------------------- */
uint32_t orc_galois_elt_from_step(int step, u64 n) {
    uint32_t m32 = (uint32_t)(2 * n);
    if (step == 0) return m32 - 1;
    int sign = step < 0;
    uint32_t pos = (uint32_t)(step < 0 ? -step : step);
    if (pos >= (n >> 1)) return 0;
    pos &= m32 - 1;
    int s = sign ? (int)(n >> 1) - (int)pos : (int)pos;
    u64 elt = 1;
    while (s--) {
        elt *= 5;
        elt &= (u64)m32 - 1;
    }
    return (uint32_t)elt;
}

void orc_galois_table(u64 n, uint32_t elt, uint32_t *table) {
    int logn = 0;
    while (((u64)1 << logn) < n) logn++;
    for (u64 i = n; i < 2 * n; i++) {
        uint32_t rev = bitrev32((uint32_t)i, logn + 1);
        u64 raw = ((u64)elt * rev) >> 1;
        raw &= n - 1;
        table[i - n] = bitrev32((uint32_t)raw, logn);
    }
}

void orc_apply_galois_ntt(const orc_ctx *c, const u64 *src, u64 *dst, int l, const uint32_t *table) {
    size_t n = c->n;
    for (int i = 0; i < l; i++)
        for (size_t j = 0; j < n; j++) dst[(size_t)i * n + j] = src[(size_t)i * n + table[j]];
}

/* apply_galois_permutation, galois.cu:20-39 (coefficient domain, BFV) */
void orc_apply_galois_coeff(const orc_ctx *c, const u64 *src, u64 *dst, int l, uint32_t elt) {
    size_t n = c->n;
    for (int i = 0; i < l; i++) {
        u64 q = c->primes[i];
        for (size_t j = 0; j < n; j++) {
            u64 idx = ((u64)j * elt) & (2 * n - 1);
            u64 v = src[(size_t)i * n + j];
            if (idx >= n) v = v ? q - v : 0;
            dst[(size_t)i * n + (idx & (n - 1))] = v;
        }
    }
}

void orc_apply_galois(const orc_ctx *c, int l, u64 *ct, uint32_t elt, const u64 *glk) {
    size_t n = c->n, poly = (size_t)l * n;
    uint32_t *table = (uint32_t *)malloc(n * 4);
    u64 *tmp = (u64 *)malloc(poly * 8);
    orc_galois_table(n, elt, table);
    if (c->scheme == ORC_SCHEME_BFV) { /* evaluate.cu:1596-1609 */
        orc_apply_galois_coeff(c, ct, tmp, l, elt);
        memcpy(ct, tmp, poly * 8);
        orc_apply_galois_coeff(c, ct + poly, tmp, l, elt);
        memset(ct + poly, 0, poly * 8);
        orc_keyswitch(c, l, ct, tmp, glk);
        free(table);
        free(tmp);
        return;
    }
    orc_apply_galois_ntt(c, ct, tmp, l, table);
    memcpy(ct, tmp, poly * 8);
    orc_apply_galois_ntt(c, ct + poly, tmp, l, table);
    memset(ct + poly, 0, poly * 8);
    orc_keyswitch(c, l, ct, tmp, glk);
    free(table);
    free(tmp);
}

/* hoisting_inplace, evaluate.cu:1670-1865: sum over elts of the rotated ciphertext with one shared mod-up and one
 * mod-down.  glk[i] = switching key of elts[i] ([dnum][2][size_QP][n]).  BFV (:1745-1747, 1808-1810): c0 takes the
 * coefficient-form automorphism, mod-up starts from and mod-down ends in coefficient form (orc_modup /
 * orc_moddown_from_ntt branch on the scheme). */
void orc_hoisting(const orc_ctx *c, int l, u64 *ct, const uint32_t *elts, int n_elts, const u64 *const *glk) {
    size_t n = c->n, poly = (size_t)l * n;
    int m = l + c->size_P, beta = orc_beta(c, l);
    size_t m_n = (size_t)m * n;
    uint32_t *table = (uint32_t *)malloc(n * 4);
    u64 *acc_c0 = (u64 *)calloc(poly, 8), *tmp = (u64 *)malloc(poly * 8);
    u64 *up = (u64 *)malloc((size_t)beta * m_n * 8), *pup = (u64 *)malloc((size_t)beta * m_n * 8);
    u64 *acc = (u64 *)calloc(2 * m_n, 8), *cx = (u64 *)malloc(2 * m_n * 8);
    u64 *res = (u64 *)malloc(poly * 8);
    orc_modup(c, l, ct + poly, up);
    for (int e = 0; e < n_elts; e++) {
        orc_galois_table(n, elts[e], table);
        if (c->scheme == ORC_SCHEME_BFV) orc_apply_galois_coeff(c, ct, tmp, l, elts[e]);
        else orc_apply_galois_ntt(c, ct, tmp, l, table);
        orc_poly_add(c, acc_c0, tmp, acc_c0, l);
        /* the same index permutation on every limb of every digit (:1775-1778); moduli do not matter here */
        for (int b = 0; b < beta * m; b++)
            for (size_t j = 0; j < n; j++) pup[(size_t)b * n + j] = up[(size_t)b * n + table[j]];
        orc_inner_prod(c, l, pup, glk[e], cx);
        for (int k = 0; k < 2; k++)
            for (int j = 0; j < m; j++) {
                u64 q = c->primes[qlp_index(c, l, j)];
                u64 *a = acc + (size_t)k * m_n + (size_t)j * n, *b = cx + (size_t)k * m_n + (size_t)j * n;
                for (size_t x = 0; x < n; x++) a[x] = addmod(a[x], b[x], q);
            }
    }
    orc_moddown_from_ntt(c, l, acc, res);
    orc_poly_add(c, acc_c0, res, ct, l);
    orc_moddown_from_ntt(c, l, acc + m_n, ct + poly);
    free(table); free(acc_c0); free(tmp); free(up); free(pup); free(acc); free(cx); free(res);
}

/* ------------------------------------------------------------------------------------------------------
 * BFV multiplication, BEHZ variant (evaluate.cu:404-548, rns.cu:386-570,1249-1517)
 * ---------------------------------------------------------------------------------------------------- */
/* get_primes (numth.cu:207-233): NTT-friendly primes of `bits` bits, descending from 2^bits */
static int get_primes_desc(u64 n, int bits, int count, u64 *out) {
    u64 factor = 2 * n, value = ((u64)1 << bits) - factor + 1, lower = (u64)1 << (bits - 1);
    int k = 0;
    while (k < count && value > lower) {
        if (orc_is_prime(value)) out[k++] = value;
        value -= factor;
    }
    return k == count ? 0 : -1;
}

/* bit length of prod(primes) (get_significant_bit_count_uint of base_Q.big_modulus) */
static int product_bits(const u64 *primes, int cnt) {
    u64 acc[64];
    int len = 1;
    memset(acc, 0, sizeof(acc));
    acc[0] = 1;
    for (int i = 0; i < cnt; i++) {
        u64 carry = 0;
        for (int k = 0; k < len; k++) {
            u128 t = (u128)acc[k] * primes[i] + carry;
            acc[k] = (u64)t;
            carry = (u64)(t >> 64);
        }
        if (carry) acc[len++] = carry;
    }
    int bits = 0;
    u64 top = acc[len - 1];
    while (top) {
        bits++;
        top >>= 1;
    }
    return (len - 1) * 64 + bits;
}

static u64 prod_mod(const u64 *base, int n, u64 p) {
    u64 r = 1 % p;
    for (int k = 0; k < n; k++) r = orc_mulmod(r, base[k] % p, p);
    return r;
}

int orc_behz_aux(const orc_ctx *c, u64 *bsk, int *nbsk) {
    int size_Q = c->size_Q, tb = 0;
    u64 t = c->t;
    while (t) {
        tb++;
        t >>= 1;
    }
    int nB = size_Q;
    if (32 + tb + product_bits(c->primes, size_Q) >= 61 * size_Q + 61) nB++; /* rns.cu:400-406 */
    u64 pr[66];
    if (get_primes_desc(c->n, 61, nB + 1, pr)) return -1;
    /* first prime = m_sk, then B (rns.cu:414-420); Bsk = B followed by m_sk */
    for (int i = 0; i < nB; i++) bsk[i] = pr[1 + i];
    bsk[nB] = pr[0];
    *nbsk = nB + 1;
    return 0;
}

/* out[3][size_Q][n] = BEHZ product of two size-2 BFV ciphertexts (coefficient form, top level) */
int orc_bfv_multiply_behz(const orc_ctx *c, const u64 *ct1, const u64 *ct2, u64 *out) {
    const size_t n = c->n;
    const int lq = c->size_Q;
    const u64 *Q = c->primes, t = c->t, mt = (u64)1 << 32;
    u64 bsk[66];
    int nbsk;
    if (orc_behz_aux(c, bsk, &nbsk)) return -1;
    const int nB = nbsk - 1;
    const u64 msk = bsk[nB];
    orc_ctx *ax = orc_create(ORC_SCHEME_BFV, n, bsk, nbsk, 0, 0); /* NTT tables of Bsk (rns.cu:435-448) */
    if (!ax) return -1;
    int idq[64], idb[66];
    for (int i = 0; i < lq; i++) idq[i] = i;
    for (int j = 0; j < nbsk; j++) idb[j] = j;
    const size_t pq = (size_t)lq * n, pb = (size_t)nbsk * n;
    u64 *eq[2], *eb[2];
    u64 *y = (u64 *)malloc(pq * 8), *mtl = (u64 *)malloc(n * 8);
    for (int s = 0; s < 2; s++) {
        const u64 *ct = s ? ct2 : ct1;
        eq[s] = (u64 *)malloc((s ? 2 : 3) * pq * 8);
        eb[s] = (u64 *)malloc((s ? 2 : 3) * pb * 8);
        for (int p = 0; p < 2; p++) { /* BEHZ_mul_1, evaluate.cu:404-438 */
            const u64 *x = ct + p * pq;
            u64 *xq = eq[s] + p * pq, *xb = eb[s] + p * pb;
            memcpy(xq, x, pq * 8);
            orc_ntt_forward(c, xq, lq, idq);
            /* fastbconv_m_tilde (rns.cu:1249-1277): y = x * (m_tilde * qhatinv) mod q */
            for (int i = 0; i < lq; i++) {
                u64 k = orc_mulmod(mt % Q[i], orc_invmod(qhat_mod(Q, lq, i, Q[i]), Q[i]), Q[i]);
                for (size_t j = 0; j < n; j++) y[(size_t)i * n + j] = orc_mulmod(x[(size_t)i * n + j], k, Q[i]);
            }
            for (int jb = 0; jb <= nbsk; jb++) {
                u64 p_out = jb < nbsk ? bsk[jb] : mt;
                u64 *dst = jb < nbsk ? xb + (size_t)jb * n : mtl;
                u64 mat[64];
                for (int i = 0; i < lq; i++) mat[i] = qhat_mod(Q, lq, i, p_out);
                for (size_t j = 0; j < n; j++) {
                    u128 acc = 0;
                    for (int i = 0; i < lq; i++) acc = (acc + (u128)y[(size_t)i * n + j] * mat[i]) % p_out;
                    dst[j] = (u64)acc;
                }
            }
            /* sm_mrq (rns.cu:1290-1339) */
            u64 nqi = (mt - orc_invmod(prod_mod(Q, lq, mt), mt)) % mt;
            for (int jb = 0; jb < nbsk; jb++) {
                u64 pj = bsk[jb], qmod = prod_mod(Q, lq, pj), imt = orc_invmod(mt % pj, pj);
                for (size_t j = 0; j < n; j++) {
                    u64 r = orc_mulmod(mtl[j], nqi, mt);
                    if (r >= mt >> 1) r += pj - mt;
                    u64 v = (u64)(((u128)r * qmod + xb[(size_t)jb * n + j]) % pj);
                    xb[(size_t)jb * n + j] = orc_mulmod(v, imt, pj);
                }
            }
            orc_ntt_forward(ax, xb, nbsk, idb);
        }
    }
    /* step 4: dyadic products in both bases (evaluate.cu:479-500), result into eq[0]/eb[0] (3 polys) */
    {
        u64 *dq = (u64 *)malloc(3 * pq * 8), *db = (u64 *)malloc(3 * pb * 8);
        orc_tensor_2x2(c, eq[0], eq[1], dq, lq);
        orc_tensor_2x2(ax, eb[0], eb[1], db, nbsk);
        memcpy(eq[0], dq, 3 * pq * 8);
        memcpy(eb[0], db, 3 * pb * 8);
        free(dq);
        free(db);
    }
    /* steps 5-6: inverse NTT and multiplication by t (evaluate.cu:520-531) */
    for (int p = 0; p < 3; p++) {
        u64 *xq = eq[0] + p * pq, *xb = eb[0] + p * pb;
        orc_ntt_inverse(c, xq, lq, idq);
        orc_ntt_inverse(ax, xb, nbsk, idb);
        for (int i = 0; i < lq; i++)
            for (size_t j = 0; j < n; j++) xq[(size_t)i * n + j] = orc_mulmod(xq[(size_t)i * n + j], t % Q[i], Q[i]);
        for (int i = 0; i < nbsk; i++)
            for (size_t j = 0; j < n; j++) xb[(size_t)i * n + j] = orc_mulmod(xb[(size_t)i * n + j], t % bsk[i], bsk[i]);
    }
    u64 *fl = (u64 *)malloc(pb * 8), *yb = (u64 *)malloc((size_t)nB * n * 8), *alpha = (u64 *)malloc(n * 8);
    for (int p = 0; p < 3; p++) {
        u64 *xq = eq[0] + p * pq, *xb = eb[0] + p * pb, *o = out + p * pq;
        /* step 7 fast_floor (rns.cu:1343-1419): (in_Bsk - FastBconv(in_q, q -> Bsk)) * Q^-1 mod Bsk */
        for (int i = 0; i < lq; i++) {
            u64 k = orc_invmod(qhat_mod(Q, lq, i, Q[i]), Q[i]);
            for (size_t j = 0; j < n; j++) y[(size_t)i * n + j] = orc_mulmod(xq[(size_t)i * n + j], k, Q[i]);
        }
        for (int jb = 0; jb < nbsk; jb++) {
            u64 pj = bsk[jb], qinv = orc_invmod(prod_mod(Q, lq, pj), pj), mat[64];
            for (int i = 0; i < lq; i++) mat[i] = qhat_mod(Q, lq, i, pj);
            for (size_t j = 0; j < n; j++) {
                u128 acc = 0;
                for (int i = 0; i < lq; i++) acc = (acc + (u128)y[(size_t)i * n + j] * mat[i]) % pj;
                fl[(size_t)jb * n + j] = orc_mulmod(submod(xb[(size_t)jb * n + j], (u64)acc, pj), qinv, pj);
            }
        }
        /* step 8 fastbconv_sk (rns.cu:1463-1517) */
        for (int i = 0; i < nB; i++) {
            u64 k = orc_invmod(qhat_mod(bsk, nB, i, bsk[i]), bsk[i]);
            for (size_t j = 0; j < n; j++) yb[(size_t)i * n + j] = orc_mulmod(fl[(size_t)i * n + j], k, bsk[i]);
        }
        {
            u64 binv = orc_invmod(prod_mod(bsk, nB, msk), msk), mat[66];
            for (int i = 0; i < nB; i++) mat[i] = qhat_mod(bsk, nB, i, msk);
            for (size_t j = 0; j < n; j++) {
                u128 acc = 0;
                for (int i = 0; i < nB; i++) acc = (acc + (u128)yb[(size_t)i * n + j] * mat[i]) % msk;
                alpha[j] = orc_mulmod(submod((u64)acc, fl[(size_t)nB * n + j], msk), binv, msk);
            }
        }
        for (int k = 0; k < lq; k++) {
            u64 qk = Q[k], Bq = prod_mod(bsk, nB, qk), mat[66];
            for (int i = 0; i < nB; i++) mat[i] = qhat_mod(bsk, nB, i, qk);
            for (size_t j = 0; j < n; j++) {
                u128 acc = 0;
                for (int i = 0; i < nB; i++) acc = (acc + (u128)yb[(size_t)i * n + j] * mat[i]) % qk;
                u64 a = alpha[j], corr; /* multiply_and_negated_add_rns_poly, polymath.cu:606-634 */
                if (a > msk >> 1) corr = orc_mulmod((msk - a) % qk, Bq, qk);
                else corr = orc_mulmod(a % qk, qk - Bq, qk);
                o[(size_t)k * n + j] = addmod((u64)acc, corr, qk);
            }
        }
    }
    free(fl); free(yb); free(alpha); free(y); free(mtl);
    for (int s = 0; s < 2; s++) {
        free(eq[s]);
        free(eb[s]);
    }
    orc_destroy(ax);
    return 0;
}

/* multiply_inplace + relinearize_inplace for BFV/BEHZ: out[2][size_Q][n] */
int orc_bfv_multiply_relin_behz(const orc_ctx *c, const u64 *ct1, const u64 *ct2, const u64 *rlk, u64 *out) {
    size_t poly = (size_t)c->size_Q * c->n;
    u64 *d = (u64 *)malloc(3 * poly * 8);
    if (orc_bfv_multiply_behz(c, ct1, ct2, d)) {
        free(d);
        return -1;
    }
    orc_keyswitch(c, c->size_Q, d, d + 2 * poly, rlk);
    memcpy(out, d, 2 * poly * 8);
    free(d);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------
 * BFV multiplication, HPS variant (mul_tech_type::hps; evaluate.cu:647-801, rns.cu:687-792,1699-1745,
 * rns_bconv.cu:248-372).  The floating-point steps follow the reference's compiled code: nvcc contracts
 * `acc += double(x) * c` into one fused multiply-add per term, in index order (checked in the SASS of
 * oracle/_ref: I2F.F64.U64 + DFMA chains), hence fma() below.
 * ---------------------------------------------------------------------------------------------------- */
/* little-endian multi-word helpers (the reference uses multiply_uint / divide_uint / modulo_uint) */
typedef struct {
    u64 w[72];
    int len;
} big_t;
static void big_mul_word(big_t *a, u64 m) {
    u64 carry = 0;
    for (int k = 0; k < a->len; k++) {
        u128 t = (u128)a->w[k] * m + carry;
        a->w[k] = (u64)t;
        carry = (u64)(t >> 64);
    }
    if (carry) a->w[a->len++] = carry;
}
static u64 big_divmod_word(big_t *a, u64 d) { /* a <- floor(a / d), returns a mod d */
    u64 rem = 0;
    for (int k = a->len - 1; k >= 0; k--) {
        u128 cur = ((u128)rem << 64) | a->w[k];
        a->w[k] = (u64)(cur / d);
        rem = (u64)(cur % d);
    }
    while (a->len > 1 && a->w[a->len - 1] == 0) a->len--;
    return rem;
}
static u64 big_mod_word(const big_t *a, u64 d) {
    u64 rem = 0;
    for (int k = a->len - 1; k >= 0; k--) rem = (u64)((((u128)rem << 64) | a->w[k]) % d);
    return rem;
}

/* get_primes_below (numth.cu:235-263) */
static int get_primes_below(u64 n, u64 upper, int count, u64 *out) {
    u64 factor = 2 * n;
    int bits = 0;
    for (u64 v = upper; v; v >>= 1) bits++;
    if (upper < factor) return -1;
    u64 value = upper - factor, lower = (u64)1 << (bits - 1);
    int k = 0;
    while (k < count && value > lower) {
        if (orc_is_prime(value)) out[k++] = value;
        value -= factor;
    }
    return k == count ? 0 : -1;
}

/* the auxiliary base R of HPS: size_Q + 1 primes below the smallest prime of Q (rns.cu:687-694) */
int orc_hps_aux(const orc_ctx *c, u64 *R, int *nR) {
    u64 qmin = c->primes[0];
    for (int i = 1; i < c->size_Q; i++)
        if (c->primes[i] < qmin) qmin = c->primes[i];
    *nR = c->size_Q + 1;
    return get_primes_below(c->n, qmin, *nR, R);
}

static u64 sat_u64(double v) { /* static_cast<uint64_t>(double) on the device: cvt.rzi.u64.f64 saturates */
    if (!(v > 0.0)) return 0;
    if (v >= 18446744073709551616.0) return ~(u64)0;
    return (u64)v;
}

/* bConv_HPS (rns_bconv.cu:354-372): exact-ish conversion with the floating-point overflow estimate */
static void bconv_hps(const u64 *ibase, int ni, const u64 *obase, int no, const u64 *in, u64 *out, size_t n) {
    u64 hinv[72], Imod[72];
    double inv[72];
    u64 *mat = (u64 *)malloc((size_t)no * ni * 8);
    for (int i = 0; i < ni; i++) {
        hinv[i] = orc_invmod(qhat_mod(ibase, ni, i, ibase[i]), ibase[i]);
        inv[i] = 1.0 / (double)ibase[i]; /* host/rns.cu:319-324 */
    }
    for (int j = 0; j < no; j++) {
        Imod[j] = prod_mod(ibase, ni, obase[j]);
        for (int i = 0; i < ni; i++) mat[(size_t)j * ni + i] = qhat_mod(ibase, ni, i, obase[j]);
    }
#pragma omp parallel for num_threads(g_threads)
    for (size_t x = 0; x < n; x++) {
        u64 y[72];
        double frac = 0.0;
        for (int i = 0; i < ni; i++) {
            y[i] = orc_mulmod(in[(size_t)i * n + x], hinv[i], ibase[i]);
            frac = fma((double)y[i], inv[i], frac);
        }
        u64 v = (u64)llround(frac);
        for (int j = 0; j < no; j++) {
            u64 p = obase[j];
            u128 acc = 0;
            for (int i = 0; i < ni; i++) acc = (acc + (u128)y[i] * mat[(size_t)j * ni + i]) % p;
            out[(size_t)j * n + x] = submod((u64)acc, orc_mulmod(v % p, Imod[j], p), p); /* alphaQModp[v][j] */
        }
    }
    free(mat);
}

/* out[3][size_Q][n] = HPS product of two size-2 BFV ciphertexts (coefficient form, top level) */
int orc_bfv_multiply_hps(const orc_ctx *c, const u64 *ct1, const u64 *ct2, u64 *out) {
    const size_t n = c->n;
    const int lq = c->size_Q;
    const u64 *Q = c->primes, t = c->t;
    u64 R[72], S[144];
    int nR;
    if (lq > 35 || orc_hps_aux(c, R, &nR)) return -1;
    const int ls = lq + nR;
    memcpy(S, Q, lq * 8);
    memcpy(S + lq, R, nR * 8);
    orc_ctx *sx = orc_create(ORC_SCHEME_BFV, n, S, ls, 0, 0); /* gpu_QlRl_tables (rns.cu:700-714) */
    if (!sx) return -1;
    int ids[144];
    for (int i = 0; i < ls; i++) ids[i] = i;
    const size_t ps = (size_t)ls * n, pq = (size_t)lq * n;
    u64 *e[2];
    for (int s = 0; s < 2; s++) {
        const u64 *ct = s ? ct2 : ct1;
        e[s] = (u64 *)malloc(3 * ps * 8);
        for (int p = 0; p < 2; p++) {
            u64 *x = e[s] + p * ps;
            memcpy(x, ct + p * pq, pq * 8);
            bconv_hps(Q, lq, R, nR, x, x + pq, n);
            orc_ntt_forward(sx, x, ls, ids);
        }
    }
    u64 *d = (u64 *)malloc(3 * ps * 8);
    orc_tensor_2x2(sx, e[0], e[1], d, ls);
    for (int p = 0; p < 3; p++) orc_ntt_inverse(sx, d + p * ps, ls, ids);

    /* scaleAndRound_HPS_QR_R tables (rns.cu:727-789): A_i = t * R * (Shat_i^-1 mod s_i) */
    double *fr = (double *)malloc(lq * sizeof(double));
    u64 *tab = (u64 *)malloc((size_t)nR * (lq + 1) * 8);
    for (int i = 0; i < ls; i++) {
        big_t A = {{1}, 1};
        for (int k = 0; k < nR; k++) big_mul_word(&A, R[k]);
        big_mul_word(&A, t);
        big_mul_word(&A, orc_invmod(qhat_mod(S, ls, i, S[i]), S[i]));
        u64 rem = big_divmod_word(&A, S[i]); /* A <- floor(A / s_i) */
        if (i < lq) {
            fr[i] = (double)rem / (double)S[i];
            for (int j = 0; j < nR; j++) tab[(size_t)j * (lq + 1) + i] = big_mod_word(&A, R[j]);
        } else {
            tab[(size_t)(i - lq) * (lq + 1) + lq] = big_mod_word(&A, R[i - lq]);
        }
    }
    u64 *tmpR = (u64 *)malloc((size_t)nR * n * 8);
    for (int p = 0; p < 3; p++) {
        const u64 *x = d + p * ps;
        /* scaleAndRound_HPS_QR_R_kernel (rns.cu:1699-1733), including alpha being re-reduced limb after limb */
#pragma omp parallel for num_threads(g_threads)
        for (size_t k = 0; k < n; k++) {
            double nu = 0.5;
            for (int i = 0; i < lq; i++) nu = fma((double)x[(size_t)i * n + k], fr[i], nu);
            const u64 alpha0 = sat_u64(nu);
            u64 alpha = alpha0;
            for (int j = 0; j < nR; j++) {
                u64 rj = R[j];
                u128 cur = 0;
                for (int i = 0; i < lq; i++) cur = (cur + (u128)x[(size_t)i * n + k] * tab[(size_t)j * (lq + 1) + i]) % rj;
                cur = (cur + (u128)x[(size_t)(lq + j) * n + k] * tab[(size_t)j * (lq + 1) + lq]) % rj;
                alpha = (g_hps_alpha_per_limb ? alpha0 : alpha) % rj;
                tmpR[(size_t)j * n + k] = addmod((u64)cur, alpha, rj);
            }
        }
        bconv_hps(R, nR, Q, lq, tmpR, out + p * pq, n);
    }
    free(tmpR); free(tab); free(fr); free(d); free(e[0]); free(e[1]);
    orc_destroy(sx);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------
 * BFV multiplication, HPS-over-Q variants (mul_tech_type::hps_overq and, with drop > 0, the arithmetic of
 * hps_overq_leveled; evaluate.cu:647-801, rns.cu:794-975,1748-1816, rns_bconv.cu:231-246, host/rns.cu:470-495).
 * Ql = the first size_Q - drop primes of Q, Rl = size_Ql primes below min(Q), QlDrop = the dropped primes.
 * ---------------------------------------------------------------------------------------------------- */
/* bConv_BEHZ_var1 (rns_bconv.cu:231-246): y_i = x_i * (-P * qhat_i^-1) mod q_i, out_j = sum_i y_i * (q_i^-1 mod p_j) */
static void bconv_var1(const u64 *ibase, int ni, const u64 *obase, int no, const u64 *in, u64 *out, size_t n) {
    u64 c1[72];
    u64 *mat = (u64 *)malloc((size_t)no * ni * 8);
    for (int i = 0; i < ni; i++) {
        u64 qi = ibase[i];
        u64 PQ = orc_mulmod(prod_mod(obase, no, qi), orc_invmod(qhat_mod(ibase, ni, i, qi), qi), qi);
        c1[i] = qi - PQ; /* host/rns.cu:481-482: q_i - (P * qhat_i^-1 mod q_i), = q_i when that product is 0 */
    }
    for (int j = 0; j < no; j++)
        for (int i = 0; i < ni; i++) mat[(size_t)j * ni + i] = orc_invmod(ibase[i] % obase[j], obase[j]);
#pragma omp parallel for num_threads(g_threads)
    for (size_t x = 0; x < n; x++) {
        u64 y[72];
        for (int i = 0; i < ni; i++) y[i] = (u64)(((u128)in[(size_t)i * n + x] * c1[i]) % ibase[i]);
        for (int j = 0; j < no; j++) {
            u64 p = obase[j];
            u128 acc = 0;
            for (int i = 0; i < ni; i++) acc = (acc + (u128)y[i] * mat[(size_t)j * ni + i]) % p;
            out[(size_t)j * n + x] = (u64)acc;
        }
    }
    free(mat);
}

/* scaleAndRound_HPS_QlRl_Ql_kernel (rns.cu:1748-1784), also used as scaleAndRound_HPS_Q_Ql (:1797-1805):
 * out_i = sum_j xb_j * tab[i][j] + xa_i * tab[i][nb] + alpha  mod a_i,  alpha = trunc(0.5 + sum_j double(xb_j) * frac_j),
 * re-reduced limb after limb like the reference does.  xa = [na][n] (base A = outputs), xb = [nb][n]. */
static void scale_round_to_a(const u64 *A, int na, int nb, const u64 *tab, const double *frac, const u64 *xa,
                             const u64 *xb, u64 *out, size_t n) {
#pragma omp parallel for num_threads(g_threads)
    for (size_t k = 0; k < n; k++) {
        double nu = 0.5;
        for (int j = 0; j < nb; j++) nu = fma((double)xb[(size_t)j * n + k], frac[j], nu);
        u64 alpha = sat_u64(nu);
        u64 xi[72];
        for (int i = 0; i < na; i++) xi[i] = xa[(size_t)i * n + k]; /* out may alias xa */
        for (int i = 0; i < na; i++) {
            u64 q = A[i];
            u128 cur = 0;
            for (int j = 0; j < nb; j++) cur = (cur + (u128)xb[(size_t)j * n + k] * tab[(size_t)i * (nb + 1) + j]) % q;
            cur = (cur + (u128)xi[i] * tab[(size_t)i * (nb + 1) + nb]) % q;
            alpha %= q;
            out[(size_t)i * n + k] = addmod((u64)cur, alpha, q);
        }
    }
}

/* tables of the two scale-and-round forms: W_i = mult * prod(num) * (Shat_i^-1 mod s_i) over S = A u B (rns.cu:826-893
 * with mult = t, num = A; rns.cu:921-968 with mult = 1, num = A, S = Q): frac_j = (W_{na+j} mod b_j) / b_j,
 * tab[i][j] = floor(W_{na+j} / b_j) mod a_i, tab[i][nb] = floor(W_i / a_i) mod a_i */
static void scale_round_tables(const u64 *A, int na, const u64 *B, int nb, u64 mult, u64 *tab, double *frac) {
    u64 S[144];
    memcpy(S, A, na * 8);
    memcpy(S + na, B, nb * 8);
    for (int i = 0; i < na + nb; i++) {
        big_t W = {{1}, 1};
        for (int k = 0; k < na; k++) big_mul_word(&W, A[k]);
        big_mul_word(&W, mult);
        big_mul_word(&W, orc_invmod(qhat_mod(S, na + nb, i, S[i]), S[i]));
        u64 rem = big_divmod_word(&W, S[i]); /* W <- floor(W / s_i) */
        if (i >= na) {
            frac[i - na] = (double)rem / (double)S[i];
            for (int a = 0; a < na; a++) tab[(size_t)a * (nb + 1) + (i - na)] = big_mod_word(&W, A[a]);
        } else {
            tab[(size_t)i * (nb + 1) + nb] = big_mod_word(&W, A[i]);
        }
    }
}

/* out[3][size_Q][n]; drop = levels dropped (0 for mul_tech hps_overq) */
int orc_bfv_multiply_hps_overq(const orc_ctx *c, const u64 *ct1, const u64 *ct2, u64 *out, int drop) {
    const size_t n = c->n;
    const int lq = c->size_Q, ll = lq - drop;
    const u64 *Q = c->primes, t = c->t;
    if (lq > 35 || drop < 0 || ll < 1) return -1;
    u64 R[72], S[144];
    u64 qmin = Q[0];
    for (int i = 1; i < lq; i++)
        if (Q[i] < qmin) qmin = Q[i];
    if (get_primes_below(n, qmin, ll, R)) return -1; /* rns.cu:796-801: Rl has as many primes as Ql */
    const int ls = 2 * ll;
    memcpy(S, Q, ll * 8);
    memcpy(S + ll, R, ll * 8);
    orc_ctx *sx = orc_create(ORC_SCHEME_BFV, n, S, ls, 0, 0);
    if (!sx) return -1;
    int ids[144];
    for (int i = 0; i < ls; i++) ids[i] = i;
    const size_t ps = (size_t)ls * n, pq = (size_t)lq * n, pl = (size_t)ll * n;
    /* Q -> Ql scale-and-round tables of the leveled form (rns.cu:921-968) */
    u64 *dtab = NULL;
    double *dfrac = NULL;
    if (drop) {
        dtab = (u64 *)malloc((size_t)ll * (drop + 1) * 8);
        dfrac = (double *)malloc(drop * sizeof(double));
        scale_round_tables(Q, ll, Q + ll, drop, 1, dtab, dfrac);
    }
    u64 *e[2];
    for (int s = 0; s < 2; s++) {
        const u64 *ct = s ? ct2 : ct1;
        e[s] = (u64 *)malloc(3 * ps * 8);
        for (int p = 0; p < 2; p++) {
            u64 *x = e[s] + p * ps;
            const u64 *src = ct + p * pq;
            if (s == 0) { /* evaluate.cu:706-721 */
                if (drop) scale_round_to_a(Q, ll, drop, dtab, dfrac, src, src + pl, x, n);
                else memcpy(x, src, pl * 8);
                bconv_hps(Q, ll, R, ll, x, x + pl, n);
            } else { /* evaluate.cu:752-758: Q (or Ql) -> Rl by the var1 conversion, then Rl -> Ql */
                bconv_var1(Q, drop ? lq : ll, R, ll, src, x + pl, n);
                bconv_hps(R, ll, Q, ll, x + pl, x, n);
            }
            orc_ntt_forward(sx, x, ls, ids);
        }
    }
    u64 *d = (u64 *)malloc(3 * ps * 8);
    orc_tensor_2x2(sx, e[0], e[1], d, ls);
    for (int p = 0; p < 3; p++) orc_ntt_inverse(sx, d + p * ps, ls, ids);
    u64 *tab = (u64 *)malloc((size_t)ll * (ll + 1) * 8);
    double *fr = (double *)malloc(ll * sizeof(double));
    scale_round_tables(Q, ll, R, ll, t, tab, fr);
    for (int p = 0; p < 3; p++) {
        const u64 *x = d + p * ps;
        u64 *o = out + p * pq;
        scale_round_to_a(Q, ll, ll, tab, fr, x, x + pl, o, n); /* scaleAndRound_HPS_QlRl_Ql, evaluate.cu:788 */
        if (drop) { /* ExpandCRTBasis_Ql_Q (rns.cu:1810-1834): times (QlDrop mod q_i), dropped limbs zero */
            for (int i = 0; i < ll; i++) {
                u64 f = prod_mod(Q + ll, drop, Q[i]);
                for (size_t k = 0; k < n; k++) o[(size_t)i * n + k] = orc_mulmod(o[(size_t)i * n + k], f, Q[i]);
            }
            memset(o + pl, 0, (pq - pl) * 8);
        }
    }
    free(fr); free(tab); free(d); free(e[0]); free(e[1]); free(dtab); free(dfrac);
    orc_destroy(sx);
    return 0;
}

/* keyswitch_inplace for hps_overq_leveled with `drop` levels dropped (eval_key_switch.cu:109-181): c2 over Q is scaled
 * down to Ql (scaleAndRound_HPS_Q_Ql; skipped when c2_low: c2 = [ll][n] already lives at Ql, the fused form
 * evaluate.cu:966-1022), switched at that level, and the result is expanded back to Q and added to ct = [2][size_Q][n] */
int orc_bfv_keyswitch_leveled(const orc_ctx *c, u64 *ct, const u64 *c2, const u64 *evk, int drop, int c2_low) {
    const size_t n = c->n;
    const int lq = c->size_Q, ll = lq - drop;
    const u64 *Q = c->primes;
    if (drop < 0 || ll < 1) return -1;
    if (drop == 0) {
        orc_keyswitch(c, lq, ct, c2, evk);
        return 0;
    }
    const size_t pl = (size_t)ll * n, pq = (size_t)lq * n;
    u64 *low = (u64 *)malloc(pl * 8);
    if (c2_low) memcpy(low, c2, pl * 8);
    else {
        u64 *tab = (u64 *)malloc((size_t)ll * (drop + 1) * 8);
        double *frac = (double *)malloc(drop * sizeof(double));
        scale_round_tables(Q, ll, Q + ll, drop, 1, tab, frac);
        scale_round_to_a(Q, ll, drop, tab, frac, c2, c2 + pl, low, n);
        free(tab); free(frac);
    }
    u64 *ks = (u64 *)calloc(2 * pl, 8);
    orc_keyswitch(c, ll, ks, low, evk);
    for (int k = 0; k < 2; k++)
        for (int i = 0; i < ll; i++) { /* ExpandCRTBasis_Ql_Q + add_to_ct (rns.cu:1810-1856) */
            u64 f = prod_mod(Q + ll, drop, Q[i]);
            for (size_t x = 0; x < n; x++) {
                u64 *o = ct + k * pq + (size_t)i * n + x;
                *o = addmod(*o, orc_mulmod(ks[k * pl + (size_t)i * n + x], f, Q[i]), Q[i]);
            }
        }
    free(ks); free(low);
    return 0;
}

/* hoisting_inplace under mul_tech hps_overq_leveled with `drop` levels dropped (evaluate.cu:1690-1701, 1731-1733,
 * 1761-1763, 1847-1862): both polynomials are scaled from Q to Ql (scaleAndRound_HPS_Q_Ql), hoisted there, and the
 * result expanded back to Q (ExpandCRTBasis_Ql_Q: times QlDrop mod q_i, dropped limbs zero).  ct = [2][size_Q][n] */
int orc_bfv_hoisting_leveled(const orc_ctx *c, u64 *ct, const uint32_t *elts, int n_elts, const u64 *const *glk, int drop) {
    const size_t n = c->n;
    const int lq = c->size_Q, ll = lq - drop;
    const u64 *Q = c->primes;
    if (drop < 0 || ll < 1) return -1;
    if (drop == 0) {
        orc_hoisting(c, lq, ct, elts, n_elts, glk);
        return 0;
    }
    const size_t pl = (size_t)ll * n, pq = (size_t)lq * n;
    u64 *low = (u64 *)malloc(2 * pl * 8);
    u64 *tab = (u64 *)malloc((size_t)ll * (drop + 1) * 8);
    double *frac = (double *)malloc(drop * sizeof(double));
    scale_round_tables(Q, ll, Q + ll, drop, 1, tab, frac);
    for (int k = 0; k < 2; k++) scale_round_to_a(Q, ll, drop, tab, frac, ct + k * pq, ct + k * pq + pl, low + k * pl, n);
    free(tab); free(frac);
    orc_hoisting(c, ll, low, elts, n_elts, glk);
    memset(ct, 0, 2 * pq * 8);
    for (int k = 0; k < 2; k++)
        for (int i = 0; i < ll; i++) {
            u64 f = prod_mod(Q + ll, drop, Q[i]);
            for (size_t x = 0; x < n; x++)
                ct[k * pq + (size_t)i * n + x] = orc_mulmod(low[k * pl + (size_t)i * n + x], f, Q[i]);
        }
    free(low);
    return 0;
}

/* bfv_mul_relin_hps for the over-Q variants (evaluate.cu:819-1026): with levels dropped, c0 and c1 are expanded back
 * to Q, c2 stays at Ql and is switched there */
int orc_bfv_multiply_relin_hps_overq(const orc_ctx *c, const u64 *ct1, const u64 *ct2, const u64 *rlk, u64 *out,
                                     int drop) {
    const size_t n = c->n;
    const int lq = c->size_Q, ll = lq - drop;
    const size_t pq = (size_t)lq * n;
    u64 *d = (u64 *)malloc(3 * pq * 8);
    if (orc_bfv_multiply_hps_overq(c, ct1, ct2, d, drop)) {
        free(d);
        return -1;
    }
    if (drop) { /* undo the expansion of c2: its Ql residues are what the fused form switches */
        for (int i = 0; i < ll; i++) {
            u64 q = c->primes[i], finv = orc_invmod(prod_mod(c->primes + ll, drop, q), q);
            for (size_t x = 0; x < n; x++) {
                u64 *v = d + 2 * pq + (size_t)i * n + x;
                *v = orc_mulmod(*v, finv, q);
            }
        }
    }
    int rc = orc_bfv_keyswitch_leveled(c, d, d + 2 * pq, rlk, drop, 1);
    memcpy(out, d, 2 * pq * 8);
    free(d);
    return rc;
}

int orc_bfv_multiply_relin_hps(const orc_ctx *c, const u64 *ct1, const u64 *ct2, const u64 *rlk, u64 *out) {
    size_t poly = (size_t)c->size_Q * c->n;
    u64 *d = (u64 *)malloc(3 * poly * 8);
    if (orc_bfv_multiply_hps(c, ct1, ct2, d)) {
        free(d);
        return -1;
    }
    orc_keyswitch(c, c->size_Q, d, d + 2 * poly, rlk); /* bfv_mul_relin_hps tail, evaluate.cu:966-1025 */
    memcpy(out, d, 2 * poly * 8);
    free(d);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------
 * decryption (PhantomSecretKey::ckks_decrypt / bgv_decrypt / bfv_decrypt, reference src/secretkey.cu:533-691).
 * sk_pow = [size - 1][size_QP][n]: powers s, s^2, ... of the secret key in NTT form at the key level
 * (secret_key_array(), secretkey.cu:247-295); limb i of the data level uses limb i of every power.
 * ---------------------------------------------------------------------------------------------------- */
static int bits_u64(u64 v) {
    int b = 0;
    while (v) b++, v >>= 1;
    return b;
}

/* inner product c_0 + sum_i c_i s^i (CKKS/BGV: NTT form throughout; BFV: c_i to NTT form, sum back to coefficients, + c_0) */
static void decrypt_inner(const orc_ctx *c, int l, const u64 *ct, int size, const u64 *sk_pow, u64 *acc, int bfv) {
    size_t n = c->n, pl = (size_t)l * n, pk = (size_t)c->size_QP * n;
    int idx[64];
    for (int i = 0; i < l; i++) idx[i] = i;
    u64 *tmp = (u64 *)malloc(pl * 8);
    if (bfv) memset(acc, 0, pl * 8);
    else memcpy(acc, ct, pl * 8);
    for (int k = 1; k < size; k++) {
        memcpy(tmp, ct + k * pl, pl * 8);
        if (bfv) orc_ntt_forward(c, tmp, l, idx); /* secretkey.cu:603-606 */
        const u64 *sk = sk_pow + (size_t)(k - 1) * pk;
        for (int i = 0; i < l; i++) {
            u64 q = c->primes[i];
            for (size_t x = 0; x < n; x++) { /* multiply_and_add_rns_poly */
                size_t j = (size_t)i * n + x;
                acc[j] = addmod(orc_mulmod(tmp[j], sk[j], q), acc[j], q);
            }
        }
    }
    if (bfv) {
        orc_ntt_inverse(c, acc, l, idx);
        for (int i = 0; i < l; i++)
            for (size_t x = 0; x < n; x++) acc[(size_t)i * n + x] = addmod(ct[(size_t)i * n + x], acc[(size_t)i * n + x], c->primes[i]);
    }
    free(tmp);
}

/* hps_decrypt_scale_and_round (rns.cu:1519-1692, tables :594-680): round(t/Q * x) mod t from the RNS residues, with the
 * reference's four kernel variants selected by the bit budgets; every `sum += double(a) * c` is one fma, in index order */
static void hps_decrypt_scale_round(const orc_ctx *c, int l, const u64 *x, u64 *out) {
    const size_t n = c->n;
    const u64 *Q = c->primes, t = c->t;
    u64 qmax = 0;
    for (int i = 0; i < c->size_Q; i++)
        if (Q[i] > qmax) qmax = Q[i];
    const int qMSB = bits_u64(qmax), sizeQMSB = bits_u64((u64)l), tMSB = bits_u64(t), hf = qMSB >> 1;
    u64 mt[64], mtB[64];
    double fr[64], frB[64];
    for (int i = 0; i < l; i++) {
        u64 qi = Q[i], hinv = orc_invmod(qhat_mod(Q, l, i, qi), qi);
        u128 w = (u128)t * hinv;
        mt[i] = (u64)((w / qi) % t);
        fr[i] = (double)(u64)(w % qi) / (double)qi;
        u64 hb = (u64)(((u128)hinv << hf) % qi);
        w = (u128)t * hb;
        mtB[i] = (u64)((w / qi) % t);
        frB[i] = (double)(u64)(w % qi) / (double)qi;
    }
    const int large = qMSB + sizeQMSB >= 52;
    const int lazy = large ? (hf + tMSB + sizeQMSB) < 52 : (qMSB + tMSB + sizeQMSB) < 52;
    const double tInv = 1. / (double)t;
#pragma omp parallel for num_threads(g_threads)
    for (size_t k = 0; k < n; k++) {
        double fs = 0.0;
        u64 is = 0;
        for (int i = 0; i < l; i++) {
            u64 v = x[(size_t)i * n + k];
            if (!large) {
                fs = fma((double)v, fr[i], fs);
                is += lazy ? v * mt[i] : orc_mulmod(v, mt[i], t);
            } else {
                u64 hi = v >> hf, lo = v & (((u64)1 << hf) - 1);
                fs = fma((double)lo, fr[i], fs);
                fs = fma((double)hi, frB[i], fs);
                is += lazy ? lo * mt[i] : orc_mulmod(lo, mt[i], t);
                is += lazy ? hi * mtB[i] : orc_mulmod(hi, mtB[i], t);
            }
        }
        fs += (double)is;
        u64 quot = sat_u64(fs * tInv);
        fs -= (double)(t * quot);
        out[k] = (u64)llround(fs);
    }
}

/* behz_decrypt_scale_and_round (rns.cu:1008-1080, constants :331-390): gamma = the largest 61-bit NTT prime */
static int behz_decrypt_scale_round(const orc_ctx *c, int l, const u64 *x, u64 *out) {
    const size_t n = c->n;
    const u64 *Q = c->primes, t = c->t;
    u64 gamma;
    {
        int b61 = 61;
        if (orc_create_primes(n, &b61, 1, &gamma)) return -1;
    }
    const u64 tg[2] = {t, gamma};
    u64 *sc = (u64 *)malloc((size_t)l * n * 8), *o2 = (u64 *)malloc(2 * n * 8);
    for (int i = 0; i < l; i++) {
        u64 f = orc_mulmod(t % Q[i], gamma % Q[i], Q[i]);
        for (size_t k = 0; k < n; k++) sc[(size_t)i * n + k] = orc_mulmod(x[(size_t)i * n + k], f, Q[i]);
    }
    const u64 *in[64];
    u64 *op[2] = {o2, o2 + n};
    for (int i = 0; i < l; i++) in[i] = sc + (size_t)i * n;
    bconv(Q, l, tg, 2, in, op, n, 0); /* base_q_to_t_gamma_conv_.bConv_BEHZ */
    u64 ninv[2];
    for (int j = 0; j < 2; j++) {
        u64 qm = prod_mod(Q, l, tg[j]);
        ninv[j] = (tg[j] - orc_invmod(qm, tg[j])) % tg[j];
    }
    const u64 inv_gamma_t = orc_invmod(gamma % t, t), half = gamma >> 1;
    for (size_t k = 0; k < n; k++) {
        u64 a = orc_mulmod(o2[k], ninv[0], t), g = orc_mulmod(o2[n + k], ninv[1], gamma), tmp;
        if (g > half) tmp = addmod(a, (gamma - g) % t, t); /* perform_final_multiplication, :984-1006 */
        else tmp = submod(a, g % t, t);
        out[k] = orc_mulmod(tmp, inv_gamma_t, t);
    }
    free(sc); free(o2);
    return 0;
}

/* exact_convert_array Q_l -> t (rns_bconv.cu:374-430; bgv decrypt_mod_t, rns.cu:1237-1240): v accumulates IEEE
 * quotients double(y_i) / double(q_i) in index order, rounded half away from zero */
static void exact_convert_to_t(const orc_ctx *c, int l, const u64 *x, u64 *out) {
    const size_t n = c->n;
    const u64 *Q = c->primes, t = c->t;
    u64 hinv[64], mat[64];
    for (int i = 0; i < l; i++) hinv[i] = orc_invmod(qhat_mod(Q, l, i, Q[i]), Q[i]), mat[i] = qhat_mod(Q, l, i, t);
    const u64 q_mod_t = prod_mod(Q, l, t);
#pragma omp parallel for num_threads(g_threads)
    for (size_t k = 0; k < n; k++) {
        double v = 0.0;
        u128 ip = 0;
        for (int i = 0; i < l; i++) {
            u64 y = orc_mulmod(x[(size_t)i * n + k], hinv[i], Q[i]);
            ip = (ip + (u128)y * mat[i]) % t;
            v += (double)y / (double)Q[i];
        }
        u64 rv = (u64)round(v);
        out[k] = submod((u64)ip, orc_mulmod(rv % t, q_mod_t, t), t);
    }
}

/* out: CKKS [l][n] (NTT form); BGV / BFV [n] residues mod t.  mul_tech as host/encryptionparams.h:25-35 (BFV only);
 * correction_factor: BGV (ciphertext.h), 1 otherwise */
int orc_decrypt(const orc_ctx *c, int l, const u64 *ct, int size, const u64 *sk_pow, int mul_tech, u64 correction_factor,
                u64 *out) {
    const size_t n = c->n;
    if (size < 1 || l < 1 || l > 64) return -1;
    if (c->scheme == ORC_SCHEME_CKKS) {
        decrypt_inner(c, l, ct, size, sk_pow, out, 0);
        return 0;
    }
    u64 *acc = (u64 *)malloc((size_t)l * n * 8);
    int rc = 0;
    if (c->scheme == ORC_SCHEME_BGV) {
        int idx[64];
        for (int i = 0; i < l; i++) idx[i] = i;
        decrypt_inner(c, l, ct, size, sk_pow, acc, 0);
        orc_ntt_inverse(c, acc, l, idx);
        exact_convert_to_t(c, l, acc, out);
        if (correction_factor != 1) { /* secretkey.cu:681-690 */
            u64 fix = orc_invmod(correction_factor % c->t, c->t);
            for (size_t k = 0; k < n; k++) out[k] = orc_mulmod(out[k], fix, c->t);
        }
    } else {
        decrypt_inner(c, l, ct, size, sk_pow, acc, 1);
        if (mul_tech == 1) rc = behz_decrypt_scale_round(c, l, acc, out);
        else hps_decrypt_scale_round(c, l, acc, out);
    }
    free(acc);
    return rc;
}

/* ------------------------------------------------------------------------------------------------------
 * BFV / BGV batch encoder (PhantomBatchEncoder, reference src/batchencoder.cu): slot i lives at coefficient position
 * index_map[i] of the NTT-form vector mod t (matrix representation: two rows of N/2 slots, generator 5), the plaintext
 * is its inverse negacyclic NTT mod t.  t must be a prime = 1 mod 2N.
 * ---------------------------------------------------------------------------------------------------- */
static void batch_index_map(u64 n, u64 *map) { /* populate_matrix_reps_index_map, batchencoder.cu:26-49 */
    int logn = 0;
    while (((u64)1 << logn) < n) logn++;
    u64 row = n >> 1, m = n << 1, pos = 1;
    for (u64 i = 0; i < row; i++) {
        map[i] = bitrev32((uint32_t)((pos - 1) >> 1), logn);
        map[row | i] = bitrev32((uint32_t)((m - pos - 1) >> 1), logn);
        pos = (pos * 5) & (m - 1);
    }
}
int orc_batch_encode(u64 n, u64 t, const u64 *values, u64 count, u64 *plain) {
    if (count > n) return -1;
    orc_ctx *c = orc_create(ORC_SCHEME_CKKS, n, &t, 1, 0, 0); /* NTT tables mod t (gpu_plain_tables) */
    if (!c) return -1;
    u64 *map = (u64 *)malloc(n * 8);
    batch_index_map(n, map);
    for (u64 i = 0; i < n; i++) {
        u64 v = i < count ? values[i] : 0;
        plain[map[i]] = i < count ? v + (v >> 63) * t : 0; /* encode_gpu, batchencoder.cu:51-60 */
    }
    int zero = 0;
    orc_ntt_inverse(c, plain, 1, &zero);
    free(map);
    orc_destroy(c);
    return 0;
}
int orc_batch_decode(u64 n, u64 t, const u64 *plain, u64 *values) {
    orc_ctx *c = orc_create(ORC_SCHEME_CKKS, n, &t, 1, 0, 0);
    if (!c) return -1;
    u64 *map = (u64 *)malloc(n * 8), *tmp = (u64 *)malloc(n * 8);
    batch_index_map(n, map);
    memcpy(tmp, plain, n * 8);
    int zero = 0;
    orc_ntt_forward(c, tmp, 1, &zero);
    for (u64 i = 0; i < n; i++) values[i] = tmp[map[i]]; /* decode_gpu, batchencoder.cu:91-95 */
    free(map); free(tmp);
    orc_destroy(c);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------
 * CKKS encoder, encoding direction (PhantomCKKSEncoder::encode_internal, reference src/ckks.cu:66-135; special inverse
 * FFT src/fft.cu:219-345,386-422; roots src/fft.cu:13-43; decompose_array src/rns_base.cu:49-103,155-173).
 * Floating point: every operation below is the one the reference's compiled kernels execute, in the same order -- the
 * Gentleman-Sande butterfly is (x0 + x1, (x0 - x1) * w) with the complex product contracted by nvcc to
 * re = fma(d.x, w.x, -(d.y * w.y)) and one of two forms of im (see the loop; SASS of oracle/_ref), then * scalar, round().
 * ---------------------------------------------------------------------------------------------------- */
static void ckks_root(u64 m, u64 index, double *re, double *im) { /* ComplexRoots::get_root, 8-fold symmetry */
    const double PI_ = 3.1415926535897932384626433832795028842;
    index &= m - 1;
    if (index <= m / 8) {
        double th = 2 * PI_ * (double)index / (double)m;
        *re = cos(th), *im = sin(th); /* std::polar(1.0, theta) */
    } else if (index <= m / 4) {
        double a, b;
        ckks_root(m, m / 4 - index, &a, &b);
        *re = b, *im = a;
    } else if (index <= m / 2) {
        double a, b;
        ckks_root(m, m / 2 - index, &a, &b);
        *re = 0.0 - a, *im = 0.0 - (-b); /* cuCsub({0,0}, cuConj(r)) */
    } else if (index <= 3 * m / 4) {
        double a, b;
        ckks_root(m, index - m / 2, &a, &b);
        *re = 0.0 - a, *im = 0.0 - b;
    } else {
        double a, b;
        ckks_root(m, m - index, &a, &b);
        *re = a, *im = -b;
    }
}

/* values = count <= n/2 complex numbers (re, im interleaved); out = [l][n] residues in NTT form.
 * Returns 0, -1 on bad arguments, -2 when the encoded values are too large (ckks.cu:122-124) or need the reference's slow
 * multi-word path (more than 128 bits), which this restatement does not cover. */
int orc_ckks_encode(const orc_ctx *c, int l, const double *values, u64 count, double scale, u64 *out) {
    const u64 n = c->n, slots = n >> 1, m = n << 1;
    if (count == 0 || count > slots || !(scale > 0)) return -1;
    int logs = 0;
    while (((u64)1 << logs) < slots) logs++;
    double *xr = (double *)calloc(slots, sizeof(double)), *xi = (double *)calloc(slots, sizeof(double));
    for (u64 i = 0; i < count; i++) { /* bit_reverse_kernel, ckks.cu:9-15 */
        u64 r = logs ? bitrev32((uint32_t)i, logs) : 0;
        xr[r] = values[2 * i], xi[r] = values[2 * i + 1];
    }
    u64 *group = (u64 *)malloc((slots / 2 + 1) * 8);
    {
        u64 pos = 1;
        for (u64 i = 0; i < slots / 2; i++) group[i] = pos, pos = (pos * 5) & (m - 1);
    }
    const double fix = scale / (double)slots;
    for (int iter = logs - 1; iter >= 0; iter--) { /* special_fft_backward */
        const int logPairs = logs - iter - 1;
        const u64 pairs = (u64)1 << logPairs;
#pragma omp parallel for num_threads(g_threads)
        for (u64 tid = 0; tid < slots / 2; tid++) {
            u64 k = tid >> logPairs, j = tid & (pairs - 1), a = 2 * k * pairs + j, b = a + pairs;
            uint32_t kk = (uint32_t)(k << logPairs);
            u64 gi = (u64)(bitrev32(kk, 32) >> (33 - logs)); /* __brev(k << logPairs) >> (33 - logn) */
            u64 psi = (group[gi] << logPairs) & (m - 1);
            double wr, wi;
            ckks_root(m, m - psi, &wr, &wi); /* twiddles[M - psiIdx], the table holds get_root(i) for i < M */
            double sr = xr[a] + xr[b], si = xi[a] + xi[b], dr = xr[a] - xr[b], di = xi[a] - xi[b];
            double t1 = di * wi;
            xr[a] = sr, xi[a] = si;
            xr[b] = fma(dr, wr, -t1);
            /* the imaginary part was contracted differently in the reference's two kernels (SASS of oracle/_ref): the
             * shared-memory kernel, which runs the stages with pairs <= SWITCH_POINT / 2 = 1024 (fft.cu:219-282,396-412),
             * has fma(d.y, w.x, d.x * w.y); the one-stage kernel (fft.cu:296-345) fma(d.x, w.y, d.y * w.x) */
            xi[b] = logPairs <= 10 ? fma(di, wr, dr * wi) : fma(dr, wi, di * wr);
            if (iter == 0) {
                xr[a] *= fix, xi[a] *= fix, xr[b] *= fix, xi[b] *= fix;
            }
        }
    }
    if (logs == 0) xr[0] *= fix, xi[0] *= fix; /* n = 2: no stage; not reachable for supported degrees */
    double mx = 0;
    for (u64 i = 0; i < slots; i++) mx = fmax(mx, fmax(fabs(xr[i]), fabs(xi[i])));
    int bits = (int)ceil(log2(fmax(mx, 1.0))) + 1, qbits = 0;
    {
        big_t Qb = {{1}, 1};
        for (int i = 0; i < l; i++) big_mul_word(&Qb, c->primes[i]);
        qbits = 64 * (Qb.len - 1);
        for (u64 v = Qb.w[Qb.len - 1]; v; v >>= 1) qbits++;
    }
    int rc = 0;
    if (bits >= qbits || bits > 128) rc = -2;
    for (int i = 0; i < l && !rc; i++) {
        const u64 q = c->primes[i];
        for (u64 x = 0; x < n; x++) {
            double cd = round(x < slots ? xr[x] : xi[x - slots]);
            int negv = signbit(cd) ? 1 : 0;
            u64 r;
            if (bits <= 64) r = sat_u64(fabs(cd)) % q;
            else {
                double ad = fabs(cd);
                u128 v = ((u128)sat_u64(ad / 18446744073709551616.0) << 64) | sat_u64(fmod(ad, 18446744073709551616.0));
                r = (u64)(v % q);
            }
            out[(size_t)i * n + x] = negv ? q - r : r; /* a negative zero gives q, as in the reference */
        }
    }
    if (!rc) {
        int idx[64];
        for (int i = 0; i < l; i++) idx[i] = i;
        orc_ntt_forward(c, out, l, idx);
    }
    free(group); free(xr); free(xi);
    return rc;
}

/* PhantomCKKSEncoder::decode_internal (src/ckks.cu:137-190): inverse NTT, CRT composition to multi-word integers
 * (compose_array_kernel, src/rns_base.cu:174-244), centring against (Q + 1) / 2, conversion to double word by word
 * (plain multiply then add: the ternary in the reference keeps nvcc from fusing them), forward special FFT
 * (src/fft.cu:90-218,352-384; imaginary part of the product: fma(x1.y, w.x, x1.x * w.y) in the shared-memory kernel, i.e.
 * for pairs <= 1024, fma(x1.x, w.y, x1.y * w.x) in the one-stage kernel), bit-reversed readout.
 * plain = [l][n] in NTT form; out = n/2 complex values (re, im interleaved). */
int orc_ckks_decode(const orc_ctx *c, int l, const u64 *plain, double scale, double *out) {
    const u64 n = c->n, slots = n >> 1, m = n << 1;
    if (!(scale > 0) || l < 1 || l > 64) return -1;
    {   /* "scale out of bounds", ckks.cu:148-151 */
        big_t Qb = {{1}, 1};
        for (int i = 0; i < l; i++) big_mul_word(&Qb, c->primes[i]);
        int qbits = 64 * (Qb.len - 1);
        for (u64 v = Qb.w[Qb.len - 1]; v; v >>= 1) qbits++;
        if ((int)log2(scale) >= qbits) return -2;
    }
    int logs = 0;
    while (((u64)1 << logs) < slots) logs++;
    const u64 *Q = c->primes;
    u64 *w = (u64 *)malloc((size_t)l * n * 8);
    memcpy(w, plain, (size_t)l * n * 8);
    int idx[64];
    for (int i = 0; i < l; i++) idx[i] = i;
    orc_ntt_inverse(c, w, l, idx);
    /* multi-word constants: Q, punctured products, threshold */
    u64 Qw[64] = {0}, thr[64] = {0}, hat[64][64], hinv[64];
    {
        big_t b = {{1}, 1};
        for (int i = 0; i < l; i++) big_mul_word(&b, Q[i]);
        for (int k = 0; k < l; k++) Qw[k] = k < b.len ? b.w[k] : 0;
        u64 carry = 1; /* (Q + 1) >> 1 */
        u64 t[64];
        for (int k = 0; k < l; k++) {
            t[k] = Qw[k] + carry;
            carry = (carry && t[k] == 0) ? 1 : 0;
        }
        for (int k = 0; k < l; k++) thr[k] = (t[k] >> 1) | (k + 1 < l ? t[k + 1] << 63 : 0);
        for (int i = 0; i < l; i++) {
            big_t h = {{1}, 1};
            for (int j = 0; j < l; j++)
                if (j != i) big_mul_word(&h, Q[j]);
            for (int k = 0; k < l; k++) hat[i][k] = k < h.len ? h.w[k] : 0;
            hinv[i] = orc_invmod(qhat_mod(Q, l, i, Q[i]), Q[i]);
        }
    }
    double *xr = (double *)calloc(slots, sizeof(double)), *xi = (double *)calloc(slots, sizeof(double));
    const double inv_scale = 1.0 / scale;
#pragma omp parallel for num_threads(g_threads)
    for (u64 x = 0; x < n; x++) {
        u64 acc[64] = {0};
        if (l > 1) {
            for (int i = 0; i < l; i++) {
                u64 prod = orc_mulmod(w[(size_t)i * n + x], hinv[i], Q[i]);
                u64 tmp[65], carry = 0;
                for (int k = 0; k < l; k++) { /* hat_i * prod, low l words (multiply_uint_uint64) */
                    u128 t = (u128)hat[i][k] * prod + carry;
                    tmp[k] = (u64)t, carry = (u64)(t >> 64);
                }
                /* acc = (acc + tmp) mod Q: both < Q (add_uint_uint_mod) */
                u64 cy = 0, sum[64];
                for (int k = 0; k < l; k++) {
                    u128 t = (u128)acc[k] + tmp[k] + cy;
                    sum[k] = (u64)t, cy = (u64)(t >> 64);
                }
                int ge = cy != 0;
                if (!ge) {
                    ge = 1;
                    for (int k = l - 1; k >= 0; k--)
                        if (sum[k] != Qw[k]) {
                            ge = sum[k] > Qw[k];
                            break;
                        }
                }
                if (ge) {
                    u64 bw = 0;
                    for (int k = 0; k < l; k++) {
                        u128 t = (u128)sum[k] - Qw[k] - bw;
                        sum[k] = (u64)t, bw = (u64)(t >> 64) & 1;
                    }
                }
                memcpy(acc, sum, l * 8);
            }
        } else acc[0] = w[x];
        int upper = 1; /* acc >= threshold */
        for (int k = l - 1; k >= 0; k--)
            if (acc[k] != thr[k]) {
                upper = acc[k] > thr[k];
                break;
            }
        double res = 0.0, s2 = inv_scale;
        for (int k = 0; k < l; k++, s2 *= 18446744073709551616.0) {
            if (upper) {
                if (acc[k] > Qw[k]) {
                    u64 d = acc[k] - Qw[k];
                    res += d ? (double)d * s2 : 0.0;
                } else {
                    u64 d = Qw[k] - acc[k];
                    res -= d ? (double)d * s2 : 0.0;
                }
            } else {
                u64 d = acc[k];
                res += d ? (double)d * s2 : 0.0;
            }
        }
        if (x < slots) xr[x] = res;
        else xi[x - slots] = res;
    }
    u64 *group = (u64 *)malloc((slots / 2 + 1) * 8);
    {
        u64 pos = 1;
        for (u64 i = 0; i < slots / 2; i++) group[i] = pos, pos = (pos * 5) & (m - 1);
    }
    for (int iter = 0; iter < logs; iter++) { /* special_fft_forward */
        const int logPairs = logs - iter - 1;
        const u64 pairs = (u64)1 << logPairs;
#pragma omp parallel for num_threads(g_threads)
        for (u64 tid = 0; tid < slots / 2; tid++) {
            u64 k = tid >> logPairs, j = tid & (pairs - 1), a = 2 * k * pairs + j, b = a + pairs;
            uint32_t kk = (uint32_t)(k << logPairs);
            u64 gi = (u64)(bitrev32(kk, 32) >> (33 - logs));
            u64 psi = (group[gi] << logPairs) & (m - 1);
            double wr, wi;
            ckks_root(m, psi, &wr, &wi);
            double t1 = xi[b] * wi;
            double re = fma(xr[b], wr, -t1);
            double im = logPairs <= 10 ? fma(xi[b], wr, xr[b] * wi) : fma(xr[b], wi, xi[b] * wr);
            double ar = xr[a], ai = xi[a];
            xr[a] = ar + re, xi[a] = ai + im;
            xr[b] = ar - re, xi[b] = ai - im;
        }
    }
    for (u64 i = 0; i < slots; i++) { /* bit_reverse_kernel */
        u64 r = logs ? bitrev32((uint32_t)i, logs) : 0;
        out[2 * r] = xr[i], out[2 * r + 1] = xi[i];
    }
    free(group); free(xr); free(xi); free(w);
    return 0;
}

/* ------------------------------------------------------------------------------------------------------
 * rescale / mod switch
 * ---------------------------------------------------------------------------------------------------- */
void orc_rescale(const orc_ctx *c, int l, u64 *in, int size, u64 *out) {
    size_t n = c->n;
    int nl = l - 1;
    u64 qlast = c->primes[l - 1];
    for (int s = 0; s < size; s++) {
        u64 *ci = in + (size_t)s * l * n, *co = out + (size_t)s * nl * n;
        int last = l - 1;
        orc_ntt_inverse(c, ci + (size_t)last * n, 1, &last); /* rns.cu:1171 */
        for (int j = 0; j < nl; j++) {
            u64 q = c->primes[j];
            for (size_t x = 0; x < n; x++) co[(size_t)j * n + x] = ci[(size_t)last * n + x] % q; /* :1174 */
        }
        int idx[64];
        for (int j = 0; j < nl; j++) idx[j] = j;
        orc_ntt_forward(c, co, nl, idx); /* :1178 */
        for (int j = 0; j < nl; j++) {
            u64 q = c->primes[j];
            u64 inv = orc_invmod(qlast % q, q);
            for (size_t x = 0; x < n; x++) {
                size_t k = (size_t)j * n + x;
                co[k] = orc_mulmod(submod(ci[k], co[k], q), inv, q); /* :1154-1155 */
            }
        }
    }
}

void orc_mod_switch_drop(const orc_ctx *c, int l, const u64 *in, int size, u64 *out) {
    size_t n = c->n;
    for (int s = 0; s < size; s++)
        memcpy(out + (size_t)s * (l - 1) * n, in + (size_t)s * l * n, (size_t)(l - 1) * n * 8);
}

void orc_divide_round_q_last(const orc_ctx *c, int l, const u64 *in, int size, u64 *out) {
    size_t n = c->n;
    int nl = l - 1;
    u64 qlast = c->primes[l - 1];
    for (int s = 0; s < size; s++) {
        const u64 *ci = in + (size_t)s * l * n;
        u64 *co = out + (size_t)s * nl * n;
        for (int j = 0; j < nl; j++) {
            u64 q = c->primes[j];
            u64 inv = orc_invmod(qlast % q, q);
            for (size_t x = 0; x < n; x++) {
                u64 r = ci[(size_t)nl * n + x] % q;
                co[(size_t)j * n + x] = orc_mulmod(submod(ci[(size_t)j * n + x], r, q), inv, q);
            }
        }
    }
}

void orc_bgv_mod_switch(const orc_ctx *c, int l, u64 *in, int size, u64 *out) {
    size_t n = c->n;
    int nl = l - 1;
    u64 qlast = c->primes[l - 1], t = c->t;
    u64 inv_qlast_t = orc_invmod(qlast % t, t);
    int idx[64];
    for (int j = 0; j < l; j++) idx[j] = j;
    for (int s = 0; s < size; s++) {
        u64 *ci = in + (size_t)s * l * n, *co = out + (size_t)s * nl * n;
        orc_ntt_inverse(c, ci, l, idx);
        const u64 *clast = ci + (size_t)nl * n;
        for (int j = 0; j < nl; j++) {
            u64 q = c->primes[j];
            u64 inv = orc_invmod(qlast % q, q);
            u64 qlast_q = qlast % q;
            for (size_t x = 0; x < n; x++) {
                u64 v = clast[x];
                u64 delta = v % q;
                u64 tmp = orc_mulmod(v % t, inv_qlast_t, t);
                u64 corr = orc_mulmod(tmp, qlast_q, q);
                u64 r = submod(ci[(size_t)j * n + x], delta, q);
                r = addmod(r, corr, q);
                co[(size_t)j * n + x] = orc_mulmod(r, inv, q);
            }
        }
        orc_ntt_forward(c, co, nl, idx);
    }
}

/* =====================================================================================================
 * Samplers, key generation, encryption (src/prng.cu, src/secretkey.cu, src/scalingvariant.cu)
 * ===================================================================================================== */

/* salsa20_gpu, src/prng.cu:17-133, for outlen = 64 (the only length the samplers ask for): the Salsa20 core, ten double
 * rounds, over the state {key bytes 0..31 as 8 little-endian words, nonce low, nonce high, key bytes 32..55 as 6 words},
 * input added back; the 64 output bytes are the 16 words in little-endian order. */
void orc_prng_block(uint8_t out[64], const uint8_t *key, u64 nonce) {
    static const unsigned char quarter[8][4] = {{0, 4, 8, 12}, {5, 9, 13, 1}, {10, 14, 2, 6}, {15, 3, 7, 11},
                                                {0, 1, 2, 3},  {5, 6, 7, 4},  {10, 11, 8, 9}, {15, 12, 13, 14}};
    static const int rot[4] = {7, 9, 13, 18};
    uint32_t in[16], x[16];
    for (int i = 0; i < 14; i++) {
        const uint8_t *p = key + 4 * i;
        uint32_t w = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
        in[i < 8 ? i : i + 2] = w;
    }
    in[8] = (uint32_t)nonce;
    in[9] = (uint32_t)(nonce >> 32);
    memcpy(x, in, sizeof(x));
    for (int round = 0; round < 20; round += 2)
        for (int qi = 0; qi < 8; qi++) {
            const unsigned char *q = quarter[qi]; /* (a, b, c, d): b ^= (a+d)<<<7, c ^= (b+a)<<<9, d ^= (c+b)<<<13, a ^= (d+c)<<<18 */
            int tgt[4] = {q[1], q[2], q[3], q[0]}, s0[4] = {q[0], q[1], q[2], q[3]}, s1[4] = {q[3], q[0], q[1], q[2]};
            for (int k = 0; k < 4; k++) {
                uint32_t v = x[s0[k]] + x[s1[k]];
                x[tgt[k]] ^= (v << rot[k]) | (v >> (32 - rot[k]));
            }
        }
    for (int i = 0; i < 16; i++) {
        uint32_t w = x[i] + in[i];
        out[4 * i] = (uint8_t)w, out[4 * i + 1] = (uint8_t)(w >> 8), out[4 * i + 2] = (uint8_t)(w >> 16), out[4 * i + 3] = (uint8_t)(w >> 24);
    }
}

static int popcount8(unsigned v) { return __builtin_popcount(v & 0xFF); }

/* sample_ternary_poly (:142-164), sample_error_poly (:222-244), sample_uniform_poly (:174-204); the loops run over the
 * reference's thread index so that the nonces read as they do there.  out = [limbs][n] over primes 0 .. limbs-1 */
int orc_sample_poly(const orc_ctx *c, int kind, int limbs, const uint8_t *seed, u64 *out) {
    const u64 n = c->n;
    if (limbs < 1 || limbs > c->size_QP) return -1;
    uint8_t blk[64];
    if (kind == 0 || kind == 1) {
        for (u64 tid = 0; tid < n * (u64)limbs; tid++) {
            const u64 q = c->primes[tid / n];
            orc_prng_block(blk, seed, tid % n);
            if (kind == 0) {
                unsigned r = blk[0] % 3;
                out[tid] = r == 0 ? q - 1 : r - 1;
            } else {
                int cbd = popcount8(blk[0]) + popcount8(blk[1]) + popcount8(blk[2] & 0x1F) - popcount8(blk[3]) - popcount8(blk[4]) -
                          popcount8(blk[5] & 0x1F);
                out[tid] = cbd < 0 ? q - (u64)(-cbd) : (u64)cbd;
            }
        }
        return 0;
    }
    if (kind != 2) return -1;
    const u64 groups = n >> 3;
    for (u64 tid = 0; tid < groups * (u64)limbs; tid++) {
        const u64 q = c->primes[tid / groups];
        const u64 max_multiple = UINT64_MAX - (UINT64_MAX % q) - 1;
        u64 tries = 0;
        orc_prng_block(blk, seed, tid);
        tries++;
        for (int index = 0; index < 8; index++) {
            u64 r;
            for (;;) {
                memcpy(&r, blk + 8 * index, 8); /* little-endian host, like the device */
                if (r <= max_multiple) break;
                orc_prng_block(blk, seed, tid + tries * n * (u64)limbs);
                tries++;
            }
            out[tid * 8 + index] = r % q;
        }
    }
    return 0;
}

/* gen_secretkey, secretkey.cu:345-378 */
void orc_gen_secretkey(const orc_ctx *c, const uint8_t *seed, u64 *sk) {
    int idx[64];
    for (int i = 0; i < c->size_QP; i++) idx[i] = i;
    orc_sample_poly(c, 0, c->size_QP, seed, sk);
    orc_ntt_forward(c, sk, c->size_QP, idx);
}

/* encrypt_zero_symmetric, secretkey.cu:232-295.  chain_index 0: key level (size_QP limbs); data levels: BFV ciphertexts in
 * coefficient form, the others in NTT form.  ct = [2][limbs][n] */
int orc_encrypt_zero_symmetric(const orc_ctx *c, int chain_index, const u64 *sk, const uint8_t *seed_a, const uint8_t *seed_e,
                               u64 *ct) {
    if (chain_index < 0 || chain_index > c->size_Q) return -1;
    const int limbs = chain_index == 0 ? c->size_QP : c->size_Q - (chain_index - 1);
    const int ntt_form = chain_index == 0 || c->scheme != ORC_SCHEME_BFV;
    const u64 n = c->n;
    u64 *c0 = ct, *c1 = ct + (size_t)limbs * n;
    u64 *e = (u64 *)malloc((size_t)limbs * n * 8);
    int idx[64];
    for (int i = 0; i < limbs; i++) idx[i] = i;
    orc_sample_poly(c, 1, limbs, seed_e, e);
    orc_sample_poly(c, 2, limbs, seed_a, c1);
    if (ntt_form) {
        if (c->scheme == ORC_SCHEME_BGV)
            for (int i = 0; i < limbs; i++)
                for (u64 x = 0; x < n; x++) e[i * n + x] = orc_mulmod(e[i * n + x], c->t % c->primes[i], c->primes[i]);
        orc_ntt_forward(c, e, limbs, idx);
        for (int i = 0; i < limbs; i++) {
            const u64 q = c->primes[i];
            for (u64 x = 0; x < n; x++) {
                u64 v = addmod(orc_mulmod(c1[i * n + x], sk[i * n + x], q), e[i * n + x], q);
                c0[i * n + x] = v ? q - v : 0;
            }
        }
    } else {
        for (int i = 0; i < limbs; i++)
            for (u64 x = 0; x < n; x++) c0[i * n + x] = orc_mulmod(c1[i * n + x], sk[i * n + x], c->primes[i]);
        orc_ntt_inverse(c, c0, limbs, idx);
        for (int i = 0; i < limbs; i++) {
            const u64 q = c->primes[i];
            for (u64 x = 0; x < n; x++) {
                u64 v = addmod(c0[i * n + x], e[i * n + x], q);
                c0[i * n + x] = v ? q - v : 0;
            }
        }
        orc_ntt_inverse(c, c1, limbs, idx);
    }
    free(e);
    return 0;
}

/* encrypt_zero_asymmetric_internal at the first data level, secretkey.cu:10-128: (u pk_i + e) at the key level -- one error
 * polynomial for both i, the seed is not renewed between them -- then DRNSTool::moddown (rns_bconv.cu:712-761).
 * pk = [2][size_QP][n] NTT form, ct = [2][size_Q][n] */
int orc_encrypt_zero_asymmetric(const orc_ctx *c, const u64 *pk, const uint8_t *seed_u, const uint8_t *seed_e, u64 *ct) {
    const u64 n = c->n;
    const int m = c->size_QP, l = c->size_Q;
    if (c->size_P < 1) return -1;
    int idx[64];
    for (int i = 0; i < m; i++) idx[i] = i;
    u64 *u = (u64 *)malloc((size_t)m * n * 8), *e = (u64 *)malloc((size_t)m * n * 8), *cx = (u64 *)malloc((size_t)m * n * 8);
    orc_sample_poly(c, 0, m, seed_u, u);
    orc_ntt_forward(c, u, m, idx);
    for (int k = 0; k < 2; k++) {
        const u64 *pki = pk + (size_t)k * m * n;
        orc_sample_poly(c, 1, m, seed_e, e);
        if (c->scheme == ORC_SCHEME_BFV) {
            /* coefficient form: intt(pk u) + e, then the division by P without any transform (:76-92, rns_bconv.cu:744-757).
             * orc_moddown_from_ntt starts from NTT form, so hand it ntt(intt(pk u) + e) = pk u + ntt(e): the same residues */
            orc_ntt_forward(c, e, m, idx);
        } else {
            if (c->scheme == ORC_SCHEME_BGV)
                for (int i = 0; i < m; i++)
                    for (u64 x = 0; x < n; x++) e[i * n + x] = orc_mulmod(e[i * n + x], c->t % c->primes[i], c->primes[i]);
            orc_ntt_forward(c, e, m, idx);
        }
        for (int i = 0; i < m; i++) {
            const u64 q = c->primes[i];
            for (u64 x = 0; x < n; x++) cx[i * n + x] = addmod(orc_mulmod(u[i * n + x], pki[i * n + x], q), e[i * n + x], q);
        }
        orc_moddown_from_ntt(c, l, cx, ct + (size_t)k * l * n);
    }
    free(u), free(e), free(cx);
    return 0;
}

/* generate_one_kswitch_key, secretkey.cu:297-343 + multiply_temp_mod_and_add_rns_poly, polymath.cu:318-338.
 * digits = [dnum][2][size_QP][n], dnum = size_Q / size_P; seeds = dnum pairs (a, e) of 64 bytes */
int orc_gen_kswitch_key(const orc_ctx *c, const u64 *new_key, const u64 *sk, const uint8_t *seeds, u64 *digits) {
    if (c->size_P < 1 || c->size_Q % c->size_P) return -1;
    const u64 n = c->n;
    const int dnum = c->size_Q / c->size_P, alpha = c->size_P;
    const size_t digit_words = (size_t)2 * c->size_QP * n;
    for (int d = 0; d < dnum; d++)
        orc_encrypt_zero_symmetric(c, 0, sk, seeds + 128 * d, seeds + 128 * d + 64, digits + d * digit_words);
    for (int j = 0; j < dnum * alpha; j++) {
        const u64 q = c->primes[j], P = bigP_mod(c, q);
        u64 *key = digits + (size_t)(j / alpha) * digit_words + (size_t)j * n;
        for (u64 x = 0; x < n; x++) key[x] = addmod(key[x], orc_mulmod(new_key[(size_t)j * n + x], P, q), q);
    }
    return 0;
}

/* add_plain_inplace / sub_plain_inplace (evaluate.cu:1106-1224): c0 +-= plaintext.  BFV: bfv_add / bfv_sub_timesQ_overt_kernel
 * (polymath.cu:413-461), plain = [n] mod t; CKKS: add / sub_rns_poly, plain = [l][n] NTT form; BGV: plain reduced under every
 * limb, transformed, times the correction factor (multiply_scalar_and_add / _sub_rns_poly, polymath.cu:246-285) */
int orc_plain_add(const orc_ctx *c, int l, u64 *ct0, const u64 *plain, int sub, u64 correction_factor) {
    const u64 n = c->n;
    if (l < 1 || l > c->size_Q) return -1;
    if (c->scheme == ORC_SCHEME_CKKS) {
        for (int i = 0; i < l; i++)
            for (u64 x = 0; x < n; x++) {
                const u64 q = c->primes[i], a = ct0[i * n + x], b = plain[i * n + x];
                ct0[i * n + x] = sub ? submod(a, b, q) : addmod(a, b, q);
            }
        return 0;
    }
    const u64 t = c->t;
    if (c->scheme == ORC_SCHEME_BFV) {
        u64 q_mod_t = 1 % t;
        for (int i = 0; i < l; i++) q_mod_t = orc_mulmod(q_mod_t, c->primes[i] % t, t);
        const u64 neg = (t - q_mod_t) % t;
        for (int i = 0; i < l; i++) {
            const u64 q = c->primes[i], tinv = orc_invmod(t % q, q);
            for (u64 x = 0; x < n; x++) {
                const u64 v = orc_mulmod(orc_mulmod(plain[x], neg, t), tinv, q);
                ct0[i * n + x] = sub ? submod(ct0[i * n + x], v, q) : addmod(ct0[i * n + x], v, q);
            }
        }
        return 0;
    }
    u64 *lift = (u64 *)malloc((size_t)l * n * 8);
    int idx[64];
    for (int i = 0; i < l; i++) {
        idx[i] = i;
        for (u64 x = 0; x < n; x++) lift[i * n + x] = plain[x] % c->primes[i];
    }
    orc_ntt_forward(c, lift, l, idx);
    for (int i = 0; i < l; i++) {
        const u64 q = c->primes[i];
        for (u64 x = 0; x < n; x++) {
            const u64 v = orc_mulmod(lift[i * n + x], correction_factor % q, q);
            ct0[i * n + x] = sub ? submod(ct0[i * n + x], v, q) : addmod(ct0[i * n + x], v, q);
        }
    }
    free(lift);
    return 0;
}

/* the last step of encrypt_symmetric / encrypt_asymmetric (secretkey.cu:130-190, 463-530) */
int orc_encrypt_add_plain(const orc_ctx *c, int l, u64 *ct0, const u64 *plain) { return orc_plain_add(c, l, ct0, plain, 0, 1); }

/* multiply_plain_inplace (evaluate.cu:1226-1340): ct = [size][l][n].  BFV: multiply_plain_normal with abs_plain_rns_poly
 * (polymath.cu:645-664); CKKS: multiply_plain_ntt; BGV: lifted plaintext */
int orc_plain_multiply(const orc_ctx *c, int l, u64 *ct, int size, const u64 *plain) {
    const u64 n = c->n;
    if (l < 1 || l > c->size_Q || size < 1) return -1;
    int idx[64];
    for (int i = 0; i < l; i++) idx[i] = i;
    u64 *lift = NULL;
    const u64 *factor = plain;
    if (c->scheme != ORC_SCHEME_CKKS) {
        const u64 t = c->t;
        lift = (u64 *)malloc((size_t)l * n * 8);
        for (int i = 0; i < l; i++)
            for (u64 x = 0; x < n; x++) {
                u64 v = plain[x];
                if (c->scheme == ORC_SCHEME_BFV) v = v >= ((t + 1) >> 1) ? v + (c->primes[i] - t) : v;
                else v %= c->primes[i];
                lift[i * n + x] = v;
            }
        orc_ntt_forward(c, lift, l, idx);
        factor = lift;
    }
    for (int k = 0; k < size; k++) {
        u64 *ck = ct + (size_t)k * l * n;
        if (c->scheme == ORC_SCHEME_BFV) orc_ntt_forward(c, ck, l, idx);
        for (int i = 0; i < l; i++)
            for (u64 x = 0; x < n; x++) ck[i * n + x] = orc_mulmod(ck[i * n + x], factor[i * n + x], c->primes[i]);
        if (c->scheme == ORC_SCHEME_BFV) orc_ntt_inverse(c, ck, l, idx);
    }
    free(lift);
    return 0;
}

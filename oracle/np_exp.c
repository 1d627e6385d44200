/* CPU ORACLE (test infrastructure only): numpy's float64 `exp` as the reference's host evaluates it.
 *
 * The reference computes the on-target score as 1 / (1 + np.exp(x)) (/root/reference/CROPSR.py:313).
 * numpy (2.3.5 here; unpinned by the reference) dispatches float64 exp on AVX-512
 * hosts to the vendored Intel SVML routine __svml_exp8_ha, which is neither glibc's exp nor
 * correctly rounded, so only a restatement of THAT routine reproduces the reference's digits.
 * This file restates its main path (|x| < 0x1.61da04cbafe44p+9; the routine's own slow path for
 * the overflow / underflow range is not restated -- CROPSR's x lies in [-18, 9]), operation by
 * operation, from the disassembly of numpy's _multiarray_umath (constants read out of its
 * __svml_dexp_ha_data_internal_avx512 block):
 *     S  = fma_rz(x, log2e, 0x1.8000000003ff0p+48)      round toward zero: 4 fraction bits stay
 *     N  = S - shifter                                   multiple of 1/16
 *     j  = low 4 bits of S                               index of 2^(j/16)
 *     r  = fma(-N, ln2_hi, x);  r = fma(-N, ln2_lo, r)
 *     P  = R^2 (R^2 (R c6 + c5) + (R c4 + c3)) + (R c2 + c1)          (all fused)
 *     y  = fma(Th[j], fma(P, R, Tl[j]), Th[j]) * 2^floor(N)
 * tests/test_oracle_golden.py pins it against np.exp itself on hosts whose numpy takes that path.
 */
#include <fenv.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

static const uint64_t kTh[16] = {
    0x3ff0000000000000ull, 0x3ff0b5586cf9890full, 0x3ff172b83c7d517bull, 0x3ff2387a6e756238ull,
    0x3ff306fe0a31b715ull, 0x3ff3dea64c123422ull, 0x3ff4bfdad5362a27ull, 0x3ff5ab07dd485429ull,
    0x3ff6a09e667f3bcdull, 0x3ff7a11473eb0187ull, 0x3ff8ace5422aa0dbull, 0x3ff9c49182a3f090ull,
    0x3ffae89f995ad3adull, 0x3ffc199bdd85529cull, 0x3ffd5818dcfba487ull, 0x3ffea4afa2a490daull};
static const uint64_t kTl[16] = {
    0x0000000000000000ull, 0x3c979aa65d837b6dull, 0xbc801b15eaa59348ull, 0x3c968efde3a8a894ull,
    0x3c834d754db0abb6ull, 0x3c859f48a72a4c6dull, 0x3c7690cebb7aafb0ull, 0x3c9063e1e21c5409ull,
    0xbc93b3efbf5e2228ull, 0xbc7b32dcb94da51dull, 0x3c8db72fc1f0eab4ull, 0x3c71affc2b91ce27ull,
    0x3c8c1a7792cb3387ull, 0x3c736eae30af0cb3ull, 0x3c74a385a63d07a7ull, 0xbc8ff7128fd391f0ull};

static inline double u2d(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
static inline uint64_t d2u(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }

#define K_L2E    0x3ff71547652b82feull
#define K_SHIFT  0x42f8000000003ff0ull
#define K_L2H    0x3fe62e42fefa39efull
#define K_L2L    0x3c7abc9e3b39803full
#define K_RMASK  0xbfffffffffffffffull
#define K_C6     0x3f57411836940c04ull
#define K_C5     0x3f81101cbbc265c0ull
#define K_C4     0x3fa55557242d68feull
#define K_C3     0x3fc5555553939732ull
#define K_C2     0x3fe000000000d008ull
#define K_C1     0x3fefffffffffff70ull
#define K_THRESH 0x40861da04cbafe44ull

/* returns NaN for arguments the main path does not cover */
double np_exp_f64(double x) {
    if (!(fabs(x) < u2d(K_THRESH))) return NAN;
    const int mode = fegetround();
    fesetround(FE_TOWARDZERO);
    volatile double vx = x;
    const double S = fma(vx, u2d(K_L2E), u2d(K_SHIFT));
    fesetround(mode);
    const double N = S - u2d(K_SHIFT);
    const int j = (int)(d2u(S) & 15u);
    double r = fma(-N, u2d(K_L2H), x);
    r = fma(-N, u2d(K_L2L), r);
    const double R = u2d(d2u(r) & K_RMASK);
    const double R2 = R * R;
    const double a = fma(R, u2d(K_C6), u2d(K_C5));
    const double b = fma(R, u2d(K_C4), u2d(K_C3));
    const double c = fma(R, u2d(K_C2), u2d(K_C1));
    double P = fma(R2, a, b);
    P = fma(R2, P, c);
    const double t = fma(P, R, u2d(kTl[j]));
    const double y = fma(u2d(kTh[j]), t, u2d(kTh[j]));
    return ldexp(y, (int)floor(N));
}

void np_exp_f64_array(const double *x, double *y, long n) {
    for (long i = 0; i < n; ++i) y[i] = np_exp_f64(x[i]);
}

/* 1 / (1 + np.exp(x)), CROPSR.py:313 */
void np_logistic_f64_array(const double *x, double *y, long n) {
    for (long i = 0; i < n; ++i) y[i] = 1.0 / (1.0 + np_exp_f64(x[i]));
}

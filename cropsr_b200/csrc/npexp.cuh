// numpy's float64 exp on the device, bit for bit.
//
// The reference's score is 1 / (1 + np.exp(x)) (/root/reference/CROPSR.py:313).  On an AVX-512 host
// numpy evaluates float64 exp with the vendored Intel SVML routine __svml_exp8_ha -- not glibc's
// exp, not correctly rounded -- so the reference's digits are that routine's digits.  This is its
// main path restated operation by operation (same fused multiply-adds, same round-toward-zero
// range reduction, same 16-entry 2^(j/16) hi/lo tables; oracle/np_exp.c is the CPU restatement,
// pinned against np.exp itself and against tests/golden/np_exp_vectors.npz).  Arguments outside
// |x| < 707.7 (the routine's own slow path) fall back to the device exp: CROPSR's x is in [-18, 9].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

__constant__ unsigned long long c_npexp_th[16] = {
    0x3ff0000000000000ull, 0x3ff0b5586cf9890full, 0x3ff172b83c7d517bull, 0x3ff2387a6e756238ull,
    0x3ff306fe0a31b715ull, 0x3ff3dea64c123422ull, 0x3ff4bfdad5362a27ull, 0x3ff5ab07dd485429ull,
    0x3ff6a09e667f3bcdull, 0x3ff7a11473eb0187ull, 0x3ff8ace5422aa0dbull, 0x3ff9c49182a3f090ull,
    0x3ffae89f995ad3adull, 0x3ffc199bdd85529cull, 0x3ffd5818dcfba487ull, 0x3ffea4afa2a490daull};
__constant__ unsigned long long c_npexp_tl[16] = {
    0x0000000000000000ull, 0x3c979aa65d837b6dull, 0xbc801b15eaa59348ull, 0x3c968efde3a8a894ull,
    0x3c834d754db0abb6ull, 0x3c859f48a72a4c6dull, 0x3c7690cebb7aafb0ull, 0x3c9063e1e21c5409ull,
    0xbc93b3efbf5e2228ull, 0xbc7b32dcb94da51dull, 0x3c8db72fc1f0eab4ull, 0x3c71affc2b91ce27ull,
    0x3c8c1a7792cb3387ull, 0x3c736eae30af0cb3ull, 0x3c74a385a63d07a7ull, 0xbc8ff7128fd391f0ull};

__device__ __forceinline__ double np_exp_f64(double x) {
    if (!(fabs(x) < __longlong_as_double(0x40861da04cbafe44ll))) return exp(x);
    const double shifter = __longlong_as_double(0x42f8000000003ff0ll);                 // 1.5 * 2^48 + 1023 * 16 ulps
    const double S = __fma_rz(x, __longlong_as_double(0x3ff71547652b82fell), shifter);  // x * log2(e), 4 fraction bits
    const double N = __dsub_rn(S, shifter);
    const int lo = __double2loint(S);
    const int j = lo & 15;
    double r = __fma_rn(-N, __longlong_as_double(0x3fe62e42fefa39efll), x);             // ln2 hi
    r = __fma_rn(-N, __longlong_as_double(0x3c7abc9e3b39803fll), r);                    // ln2 lo
    const double R = __longlong_as_double(__double_as_longlong(r) & 0xbfffffffffffffffll);
    const double R2 = __dmul_rn(R, R);
    const double a = __fma_rn(R, __longlong_as_double(0x3f57411836940c04ll), __longlong_as_double(0x3f81101cbbc265c0ll));
    const double b = __fma_rn(R, __longlong_as_double(0x3fa55557242d68fell), __longlong_as_double(0x3fc5555553939732ll));
    const double c = __fma_rn(R, __longlong_as_double(0x3fe000000000d008ll), __longlong_as_double(0x3fefffffffffff70ll));
    double P = __fma_rn(R2, a, b);
    P = __fma_rn(R2, P, c);
    const double th = __longlong_as_double((long long)c_npexp_th[j]), tl = __longlong_as_double((long long)c_npexp_tl[j]);
    const double y = __fma_rn(th, __fma_rn(P, R, tl), th);
    // * 2^floor(N): the low mantissa bits of S are (N + 1023) * 16; y is in [1, 4) and the result is normal
    const int k = (lo >> 4) - 1023;
    return __hiloint2double(__double2hiint(y) + k * (1 << 20), __double2loint(y));
}

// 1 / (1 + np.exp(x)) as numpy evaluates it: one IEEE add, one IEEE divide
__device__ __forceinline__ double np_logistic_f64(double x) { return __ddiv_rn(1.0, __dadd_rn(1.0, np_exp_f64(x))); }

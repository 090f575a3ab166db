// exp(x) for x <= 0 in ~11 fp64 operations (libdevice's exp costs ~21 FMA slots, measured in
// profiles/r01_fp64_rates_b200.jsonl) -- the kernel-row generators evaluate N of these per candidate row on the same
// FP64 pipe the contraction needs.
//   x = (64 k + j) ln2/64 + r,  |r| <= ln2/128     (Cody-Waite split of ln2/64, n = rint(x 64/ln2) via the 2^52+2^51 trick)
//   exp(x) = 2^k * T[j] * (1 + r + r^2/2 + r^3/6 + r^4/24 + r^5/120),   T[j] = 2^(j/64)   (64-entry table)
// Truncation error r^6/720 <= 3.4e-17 relative; measured max error vs the C library over 4e6 points in [-708, 0]:
// 1.0 ulp (tools/test_fastexp.cu; libdevice exp is also a 1-ulp function).  Arguments below -700 return 0 (kernel values < 1e-304 are irrelevant here).
#pragma once

#ifndef SO_HD
#define SO_HD __host__ __device__ __forceinline__
#endif

// 2^(j/64), j = 0..63, correctly rounded (generated with 60-digit decimal arithmetic).
#define SO_EXP_TABLE_VALUES \
    0x1.0000000000000p+0, 0x1.02c9a3e778061p+0, 0x1.059b0d3158574p+0, 0x1.0874518759bc8p+0, \
    0x1.0b5586cf9890fp+0, 0x1.0e3ec32d3d1a2p+0, 0x1.11301d0125b51p+0, 0x1.1429aaea92de0p+0, \
    0x1.172b83c7d517bp+0, 0x1.1a35beb6fcb75p+0, 0x1.1d4873168b9aap+0, 0x1.2063b88628cd6p+0, \
    0x1.2387a6e756238p+0, 0x1.26b4565e27cddp+0, 0x1.29e9df51fdee1p+0, 0x1.2d285a6e4030bp+0, \
    0x1.306fe0a31b715p+0, 0x1.33c08b26416ffp+0, 0x1.371a7373aa9cbp+0, 0x1.3a7db34e59ff7p+0, \
    0x1.3dea64c123422p+0, 0x1.4160a21f72e2ap+0, 0x1.44e086061892dp+0, 0x1.486a2b5c13cd0p+0, \
    0x1.4bfdad5362a27p+0, 0x1.4f9b2769d2ca7p+0, 0x1.5342b569d4f82p+0, 0x1.56f4736b527dap+0, \
    0x1.5ab07dd485429p+0, 0x1.5e76f15ad2148p+0, 0x1.6247eb03a5585p+0, 0x1.6623882552225p+0, \
    0x1.6a09e667f3bcdp+0, 0x1.6dfb23c651a2fp+0, 0x1.71f75e8ec5f74p+0, 0x1.75feb564267c9p+0, \
    0x1.7a11473eb0187p+0, 0x1.7e2f336cf4e62p+0, 0x1.82589994cce13p+0, 0x1.868d99b4492edp+0, \
    0x1.8ace5422aa0dbp+0, 0x1.8f1ae99157736p+0, 0x1.93737b0cdc5e5p+0, 0x1.97d829fde4e50p+0, \
    0x1.9c49182a3f090p+0, 0x1.a0c667b5de565p+0, 0x1.a5503b23e255dp+0, 0x1.a9e6b5579fdbfp+0, \
    0x1.ae89f995ad3adp+0, 0x1.b33a2b84f15fbp+0, 0x1.b7f76f2fb5e47p+0, 0x1.bcc1e904bc1d2p+0, \
    0x1.c199bdd85529cp+0, 0x1.c67f12e57d14bp+0, 0x1.cb720dcef9069p+0, 0x1.d072d4a07897cp+0, \
    0x1.d5818dcfba487p+0, 0x1.da9e603db3285p+0, 0x1.dfc97337b9b5fp+0, 0x1.e502ee78b3ff6p+0, \
    0x1.ea4afa2a490dap+0, 0x1.efa1bee615a27p+0, 0x1.f50765b6e4540p+0, 0x1.fa7c1819e90d8p+0 \


SO_HD double so_exp_neg(double x, const double* __restrict__ T) {
    // branch-free: two of these per lane are interleaved by the row generators (the fp64 dependency chain, not the pipe
    // throughput, bounds them -- profiles/r01_rows_gen_stalls.md).  NaN stays NaN through the fma chain.
    const bool tiny = x < -700.0;
    x = tiny ? -700.0 : x;
    const double inv = 92.33248261689366;             // 64 / ln 2
    const double c_hi = 0x1.62e42fe000000p-7;         // ln2/64, high part (24 trailing zero bits: n*c_hi is exact)
    const double c_lo = 0x1.f473de6af278fp-36;        // ln2/64 - c_hi
    const double magic = 6755399441055744.0;          // 2^52 + 2^51
    const double tn = x * inv + magic;
    const double n = tn - magic;                      // rint(x * 64/ln2)
    double r = -n * c_hi + x;
    r = -n * c_lo + r;
#if defined(__CUDA_ARCH__)
    const int ni = __double2loint(tn);                // low word of the magic sum holds n (two's complement)
#else
    long long bits; __builtin_memcpy(&bits, &tn, 8);
    const int ni = (int)(bits & 0xffffffffLL);
#endif
    const int j = ni & 63;
    const int k = ni >> 6;                            // arithmetic shift: floor division
    double p = 8.333333333333333e-03;                 // 1/120
    p = p * r + 4.1666666666666664e-02;               // 1/24
    p = p * r + 1.6666666666666666e-01;               // 1/6
    p = p * r + 0.5;
    p = p * r + 1.0;
    p = p * r;                                        // exp(r) - 1
    const double tj = T[j];
    double y = tj * p + tj;
    // scale by 2^k: k in [-1010, 0] here, y in [1, 2): result stays normal
#if defined(__CUDA_ARCH__)
    y = __hiloint2double(__double2hiint(y) + (k << 20), __double2loint(y));
#else
    long long yb; __builtin_memcpy(&yb, &y, 8);
    yb += (long long)k * (1LL << 52);
    __builtin_memcpy(&y, &yb, 8);
#endif
    return tiny ? 0.0 : y;
}

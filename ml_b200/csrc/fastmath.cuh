// FP64 exp for non-positive arguments on the B200 FP64 pipe.
//
// The EM E-step (EM.cpp:203-207) evaluates one exp per point-component pair.  After the max-shift of the
// log-sum-exp every argument is <= 0, so the general-purpose library exp() (about 23 FP64 pipe
// instructions: range checks, a degree-11 polynomial, two-step scaling) does more than is needed.  This
// one is exp(x) = 2^e * 2^(j/32) * exp(r) with n = rint(x * 32/ln2) = 32 e + j, |r| <= ln2/64:
//   1 FMA (magic-number rounding) + 1 ADD + 2 FMA (Cody-Waite reduction) + 4 FMA + 1 MUL + 1 FMA
//   (degree-6 Taylor, truncation 2.7e-18 relative) + 1 FMA (table value) = 11 FP64 instructions,
// one 8-byte table load and integer exponent arithmetic on the ALU pipe.  Error < 1 ulp of the table
// entry + ~1 ulp of the polynomial (measured < 1.6e-16 relative against the library exp in
// tests/test_gpu_kernels.py).  exp(0) is exactly 1.  Arguments below -700 return 0 (exp(-700) is
// 1e-304 of a sum that is >= 1).
#pragma once

#include <cuda_runtime.h>

namespace mlb {

constexpr int kExpTableSize = 32;

// 2^(j/32), j = 0..31, correctly rounded (written as hex floats so host and device agree bit for bit).
__device__ __constant__ const double kExp2Table[kExpTableSize] = {
    0x1.0000000000000p+0, 0x1.059b0d3158574p+0, 0x1.0b5586cf9890fp+0, 0x1.11301d0125b51p+0,
    0x1.172b83c7d517bp+0, 0x1.1d4873168b9aap+0, 0x1.2387a6e756238p+0, 0x1.29e9df51fdee1p+0,
    0x1.306fe0a31b715p+0, 0x1.371a7373aa9cbp+0, 0x1.3dea64c123422p+0, 0x1.44e086061892dp+0,
    0x1.4bfdad5362a27p+0, 0x1.5342b569d4f82p+0, 0x1.5ab07dd485429p+0, 0x1.6247eb03a5585p+0,
    0x1.6a09e667f3bcdp+0, 0x1.71f75e8ec5f74p+0, 0x1.7a11473eb0187p+0, 0x1.82589994cce13p+0,
    0x1.8ace5422aa0dbp+0, 0x1.93737b0cdc5e5p+0, 0x1.9c49182a3f090p+0, 0x1.a5503b23e255dp+0,
    0x1.ae89f995ad3adp+0, 0x1.b7f76f2fb5e47p+0, 0x1.c199bdd85529cp+0, 0x1.cb720dcef9069p+0,
    0x1.d5818dcfba487p+0, 0x1.dfc97337b9b5fp+0, 0x1.ea4afa2a490dap+0, 0x1.f50765b6e4540p+0,
};

// Copies the table into shared memory (call with all threads of the block, then __syncthreads()).
__device__ __forceinline__ void load_exp_table(double* smem_table)
{
    if (threadIdx.x < kExpTableSize) smem_table[threadIdx.x] = kExp2Table[threadIdx.x];
}

// exp(x) for x <= 0 (also correct for small positive x up to ~700, never needed here).
// `table` is the 32-entry table, normally in shared memory.
__device__ __forceinline__ double exp_nonpositive(double x, const double* table = kExp2Table)
{
    constexpr double kMagic = 6755399441055744.0;         // 1.5 * 2^52: adding it rounds to the nearest integer
    constexpr double kInvStep = 0x1.71547652b82fep+5;      // 32 / ln 2
    constexpr double kStepHi = 0x1.62e42fee00000p-6;       // ln 2 / 32, high part (32 significant bits: fn * kStepHi is exact)
    constexpr double kStepLo = 0x1.a39ef35793c76p-38;      // ln 2 / 32 - kStepHi
    // range test on the high word (ALU pipe, not the FP64 pipe): for x <= 0, x < -700 <=> hi(x) > hi(-700)
    const bool tiny = static_cast<unsigned>(__double2hiint(x)) > 0xC085E000u;
    const double xc = tiny ? -700.0 : x;
    const double fn_magic = fma(xc, kInvStep, kMagic);
    const int n = __double2loint(fn_magic);
    const double fn = fn_magic - kMagic;
    double r = fma(fn, -kStepHi, xc);
    r = fma(fn, -kStepLo, r);
    double p = fma(r, 1.0 / 720.0, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    const double r2 = r * r;
    const double q = fma(p, r2, r);                        // exp(r) - 1
    const double t = table[n & (kExpTableSize - 1)];
    const double v = fma(t, q, t);                          // 2^(j/32) exp(r), in [0.98, 2)
    const int hi = __double2hiint(v) + ((n >> 5) << 20);   // times 2^e, e >= -1010: stays normal
    const double scaled = __hiloint2double(hi, __double2loint(v));
    return tiny ? 0.0 : scaled;
}

}  // namespace mlb

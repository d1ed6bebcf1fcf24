// Integer-preserving (fraction-free) arithmetic shared by the exact kernels: exact division
// of a 128-bit multiple of D by D through the modular inverse of D's odd part.
#pragma once

#include "xp_common.cuh"

#ifdef __CUDACC__

namespace xpx {

typedef long long i64;
typedef unsigned long long u64;
typedef __int128 i128;

__device__ __forceinline__ u64 inv_odd64(u64 d)
{ // d odd: Newton iteration for d^-1 mod 2^64
    u64 x = (d * 3) ^ 2; // 5 correct bits
    x *= 2 - d * x;
    x *= 2 - d * x;
    x *= 2 - d * x;
    x *= 2 - d * x;
    return x;
}

// Per-pivot constants of the division by D = Dodd * 2^tz.
struct Div {
    u64 D, inv; // D > 0, inv = Dodd^-1 mod 2^64
    int tz;     // trailing zeros of D, 0..62
    unsigned hs; // (64 - tz) & 63
    u64 hmask;   // tz ? ~0 : 0
    __device__ __forceinline__ void set(u64 d)
    {
        D = d;
        tz = __ffsll((i64)d) - 1;
        inv = inv_odd64(d >> tz);
        hs = (unsigned)(64 - tz) & 63u;
        hmask = tz ? ~0ull : 0ull;
    }
    __device__ __forceinline__ void set(u64 d, u64 inverse, int zeros)
    {
        D = d;
        inv = inverse;
        tz = zeros;
        hs = (unsigned)(64 - zeros) & 63u;
        hmask = zeros ? ~0ull : 0ull;
    }
};

// x / D for x an exact multiple of D, in two's complement throughout: the low 64 bits of
// x >> tz times the inverse of the odd part give the quotient mod 2^64, and the quotient fits
// int64 exactly when multiplying it back reproduces x (the low halves agree by construction,
// so only the high halves are compared).  |quotient| >= 2^63 raises `ovf` (never wraps).
__device__ __forceinline__ i64 ff_div(i128 x, const Div &dv, bool &ovf)
{
    const u64 lo = (u64)x;
    const i64 hi = (i64)(x >> 64);
    const u64 sh = (lo >> dv.tz) | (((u64)hi << dv.hs) & dv.hmask);
    const i64 q = (i64)(sh * dv.inv);
    if (__mul64hi(q, (i64)dv.D) != hi || q == (i64)0x8000000000000000ull) ovf = true;
    return q;
}

__device__ __forceinline__ i64 gcd64(i64 a, i64 b)
{
    u64 x = a < 0 ? (u64)(-a) : (u64)a, y = b < 0 ? (u64)(-b) : (u64)b;
    while (y) {
        u64 t = x % y;
        x = y;
        y = t;
    }
    return (i64)x;
}

} // namespace xpx

#endif // __CUDACC__

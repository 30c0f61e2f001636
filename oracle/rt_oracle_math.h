/* rt_oracle_math.h — TEST INFRASTRUCTURE (part of oracle/): pinned FP32 arithmetic of the oracle.
 *
 * The reference's arithmetic is HLSL compiled by an unknown driver, so the lowering of intrinsics
 * (normalize, lerp, dot, mul, reflect, sign, smoothstep) and the accuracy of log/cos/sin/exp/pow are
 * not recoverable from /root/reference.  This header PINS them:
 *   - every + - * / sqrt is one IEEE-754 binary32 operation, round-to-nearest-even, evaluated in the
 *     order the HLSL source writes it (left to right), never fused (build with -ffp-contract=off);
 *   - min/max are NaN-ignoring, as HLSL's are, and order -0 below +0 (the GPU's FMNMX; libm's fminf/fmaxf do not);
 *   - transcendentals are the polynomial routines below, built only from those IEEE operations so
 *     that any conforming FP32 machine (this CPU, the B200 with -fmad=false) returns the same bits.
 *     They follow well-known published minimax schemes (Cody–Waite reduction; fdlibm-style
 *     log/exp kernels; Cephes single-precision sin/cos kernels) and are accurate to ~1-2 ulp on the
 *     ranges this path uses, which is tighter than any GPU's native HLSL intrinsics.
 * The CUDA product has its OWN implementation of the same specification (csrc/rt_devmath.cuh);
 * the two are kept independent on purpose so that the parity tests compare two implementations.
 */
#ifndef RT_ORACLE_MATH_H
#define RT_ORACLE_MATH_H

#include <math.h>
#include <stdint.h>
#include <string.h>

namespace orc {

typedef unsigned int uint;

static inline uint  f2u(float f) { uint u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint u)  { float f; memcpy(&f, &u, 4); return f; }

static inline float rt_inf() { return u2f(0x7f800000u); }
static inline float rt_nan() { return u2f(0x7fc00000u); }

/* float -> uint as the GPU converts (cvt.rzi.u32.f32): truncation, NaN and negatives give 0, values from 2^32 up saturate.  In C++ the
 * cast of such values is undefined; HLSL does not define them either (a 1-pixel-wide image makes uv = 0 / 0). */
static inline uint f2uint_rz(float f)
{
    if (!(f > 0.0f)) return 0u;
    if (f >= 4294967296.0f) return 0xffffffffu;
    return (uint)f;
}

/* ---- HLSL scalar intrinsics ---------------------------------------------------------------------- */
/* NaN-ignoring, and -0 ordered below +0 — what the GPU's FMNMX does (libm's fminf / fmaxf return their first argument for a pair
 * of zeros, which would make the result depend on the operand order) */
static inline float min(float a, float b)
{
    if (a != a) return b;
    if (b != b) return a;
    if (a == b) return (f2u(a) & 0x80000000u) ? a : b;        /* equal: only zeros differ in bits; the negative one is the smaller */
    return a < b ? a : b;
}
static inline float max(float a, float b)
{
    if (a != a) return b;
    if (b != b) return a;
    if (a == b) return (f2u(a) & 0x80000000u) ? b : a;
    return a > b ? a : b;
}
static inline float abs(float a) { return fabsf(a); }
static inline float sqrt(float a) { return sqrtf(a); }
static inline float floor(float a) { return floorf(a); }
static inline float sign(float a) { return (float)((a > 0.0f) ? 1 : 0) - (float)((a < 0.0f) ? 1 : 0); } /* sign(0)=0, sign(NaN)=0 */
static inline float saturate(float a) { return min(max(a, 0.0f), 1.0f); }
static inline float lerp(float a, float b, float t) { return a + t * (b - a); }
static inline float smoothstep(float a, float b, float x)
{
    float t = saturate((x - a) / (b - a));
    return t * t * (3.0f - 2.0f * t);
}

/* ---- natural logarithm --------------------------------------------------------------------------- */
static inline float log(float x)
{
    uint ix = f2u(x);
    if (x != x) return x;                       /* NaN */
    if ((ix & 0x7fffffffu) == 0) return -rt_inf(); /* log(±0) = -inf */
    if (ix & 0x80000000u) return rt_nan();      /* log(negative) */
    if (ix == 0x7f800000u) return x;            /* +inf */
    int k = 0;
    if (ix < 0x00800000u) { x = x * 33554432.0f; ix = f2u(x); k = -25; } /* subnormal: scale by 2^25 */
    k += (int)(ix >> 23) - 127;
    float m = u2f((ix & 0x007fffffu) | 0x3f800000u);   /* [1,2) */
    if (m > 1.41421354f) { m = m * 0.5f; k += 1; }     /* (sqrt2/2, sqrt2] */
    float f = m - 1.0f;
    float s = f / (2.0f + f);
    float z = s * s;
    float w = z * z;
    float t1 = w * (0.40000972152f + w * 0.24279078841f);
    float t2 = z * (0.66666662693f + w * 0.28498786688f);
    float R = t2 + t1;
    float hfsq = (0.5f * f) * f;
    float dk = (float)k;
    return dk * 6.9313812256e-01f - ((hfsq - (s * (hfsq + R) + dk * 9.0580006145e-06f)) - f);
}

/* ---- exponential ------------------------------------------------------------------------------------ */
static inline float exp(float x)
{
    if (x != x) return x;
    if (x > 88.72283935546875f) return rt_inf();
    if (x < -103.972076416015625f) return 0.0f;
    float fk = x * 1.4426950216e+00f + (x < 0.0f ? -0.5f : 0.5f);
    int k = (int)fk;                                  /* truncation */
    float t = (float)k;
    float hi = x - t * 6.9314575195e-01f;
    float lo = t * 1.4286067653e-06f;
    float r = hi - lo;
    float rr = r * r;
    float c = r - rr * (1.6666625440e-1f + rr * -2.7667332906e-3f);
    float y = 1.0f - ((lo - (r * c) / (2.0f - c)) - hi);
    /* y * 2^k with two exact power-of-two factors so that the result may be subnormal */
    int k1 = k / 2, k2 = k - k1;
    return (y * u2f((uint)(k1 + 127) << 23)) * u2f((uint)(k2 + 127) << 23);
}

/* ---- sine / cosine ------------------------------------------------------------------------------------- */
static inline void sincos_reduce(float ax, int* q, float* r)
{
    int n = (int)(ax * 0.636619772367581343f + 0.5f);
    float fn = (float)n;
    float y = ((ax - fn * 1.5703125f) - fn * 4.837512969970703125e-4f) - fn * 7.549789948768648e-8f;
    *q = n & 3; *r = y;
}
static inline float sin_kernel(float r)
{
    float z = r * r;
    return ((((-1.9515295891e-4f * z + 8.3321608736e-3f) * z) - 1.6666654611e-1f) * z) * r + r;
}
static inline float cos_kernel(float r)
{
    float z = r * r;
    float y = (((2.443315711809948e-5f * z - 1.388731625493765e-3f) * z) + 4.166664568298827e-2f) * (z * z);
    return (y - 0.5f * z) + 1.0f;
}
static inline float sin(float x)
{
    float ax = fabsf(x);
    if (!(ax <= 100000.0f)) return rt_nan();          /* domain of the pinned routine; path uses [0, 2π] */
    int q; float r; sincos_reduce(ax, &q, &r);
    float v = (q & 1) ? cos_kernel(r) : sin_kernel(r);
    if (q & 2) v = -v;
    return (x < 0.0f) ? -v : v;
}
static inline float cos(float x)
{
    float ax = fabsf(x);
    if (!(ax <= 100000.0f)) return rt_nan();
    int q; float r; sincos_reduce(ax, &q, &r);
    float v = (q & 1) ? sin_kernel(r) : cos_kernel(r);
    if (q == 1 || q == 2) v = -v;
    return v;
}

/* ---- pow: HLSL lowers pow(x,y) to exp2(y*log2(x)); pinned here as exp(y*log(x)) -------------------------- */
static inline float pow(float x, float y) { return exp(y * log(x)); }

/* ---- float3 / float4 with HLSL component-wise operators -------------------------------------------------- */
struct float2 { float x, y; };
struct float3 { float x, y, z; };
struct float4 { float x, y, z, w; };

static inline float2 mk2(float x, float y) { float2 r = {x, y}; return r; }
static inline float3 mk3(float x, float y, float z) { float3 r = {x, y, z}; return r; }
static inline float3 mk3(float s) { float3 r = {s, s, s}; return r; }
static inline float3 mk3(const float* p) { float3 r = {p[0], p[1], p[2]}; return r; }
static inline float4 mk4(float3 v, float w) { float4 r = {v.x, v.y, v.z, w}; return r; }

static inline float2 operator*(float2 a, float s) { return mk2(a.x * s, a.y * s); }
static inline float2 operator/(float2 a, float s) { return mk2(a.x / s, a.y / s); }
static inline float2 operator*(float2 a, float2 b) { return mk2(a.x * b.x, a.y * b.y); }
static inline float2 operator-(float2 a, float s) { return mk2(a.x - s, a.y - s); }

static inline float3 operator+(float3 a, float3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline float3 operator-(float3 a, float3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline float3 operator*(float3 a, float3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline float3 operator/(float3 a, float3 b) { return mk3(a.x / b.x, a.y / b.y, a.z / b.z); }
static inline float3 operator*(float3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
static inline float3 operator*(float s, float3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
static inline float3 operator/(float3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
static inline float3 operator/(float s, float3 a) { return mk3(s / a.x, s / a.y, s / a.z); }
static inline float3 operator-(float3 a) { return mk3(-a.x, -a.y, -a.z); }
static inline float3& operator+=(float3& a, float3 b) { a = a + b; return a; }
static inline float3& operator*=(float3& a, float3 b) { a = a * b; return a; }
static inline float3& operator*=(float3& a, float s) { a = a * s; return a; }

static inline float dot(float3 a, float3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline float3 cross(float3 a, float3 b)
{
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
/* normalize(v) pinned as v * (1 / sqrt(dot(v,v)))  — the rsq·mul shape GPUs lower it to */
static inline float3 normalize(float3 v) { float inv = 1.0f / sqrtf(dot(v, v)); return v * inv; }
static inline float3 min(float3 a, float3 b) { return mk3(min(a.x, b.x), min(a.y, b.y), min(a.z, b.z)); }
static inline float3 max(float3 a, float3 b) { return mk3(max(a.x, b.x), max(a.y, b.y), max(a.z, b.z)); }
static inline float3 lerp(float3 a, float3 b, float t) { return a + t * (b - a); }
static inline float3 exp(float3 a) { return mk3(exp(a.x), exp(a.y), exp(a.z)); }
/* HLSL reflect(i, n) = i - 2 * n * dot(i, n); pinned as i - (2*dot(n,i))*n */
static inline float3 reflect(float3 i, float3 n) { return i - (2.0f * dot(n, i)) * n; }

/* column-major 4x4 (Unity Matrix4x4 memory order): element (row r, col c) = m[c*4 + r] */
struct float4x4 { float m[16]; };
static inline float M(const float4x4& a, int r, int c) { return a.m[c * 4 + r]; }
/* mul(M, float4 v).xyz, each row summed left to right, all four products kept (w = 0 or 1) */
static inline float3 mul_xyz(const float4x4& a, float4 v)
{
    float3 r;
    r.x = ((M(a,0,0) * v.x + M(a,0,1) * v.y) + M(a,0,2) * v.z) + M(a,0,3) * v.w;
    r.y = ((M(a,1,0) * v.x + M(a,1,1) * v.y) + M(a,1,2) * v.z) + M(a,1,3) * v.w;
    r.z = ((M(a,2,0) * v.x + M(a,2,1) * v.y) + M(a,2,2) * v.z) + M(a,2,3) * v.w;
    return r;
}

} /* namespace orc */
#endif

/*
 * vmath.h -- float3 helpers shared by host C++ and device code.
 *
 * Parity rule: every helper evaluates in the same operand order as the reference's Vec3 helpers
 * (lighter_int.hpp:352-442): dot = x*x' + y*y' + z*z' summed left to right, Normalized() multiplies
 * by 1.0f/sqrtf(len^2) and maps the zero vector to zero (lighter_int.hpp:404-415).  Device code is
 * compiled with -fmad=false and IEEE div/sqrt, host code with -ffp-contract=off, so results are
 * bit-identical to the x86-64 SSE2 build of the reference.
 */
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#  define LB_HD __host__ __device__ __forceinline__
#else
#  define LB_HD inline
#endif

#define LB_SMALL 0.001f                 /* SMALL_FLOAT, lighter_int.hpp:261 */

struct V3 { float x, y, z; };

LB_HD V3 mk3(float x, float y, float z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
LB_HD V3 mk3(float s) { return mk3(s, s, s); }
LB_HD V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
LB_HD V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
LB_HD V3 operator*(V3 a, V3 b) { return mk3(a.x * b.x, a.y * b.y, a.z * b.z); }
LB_HD V3 operator*(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
LB_HD V3 operator*(float s, V3 a) { return mk3(s * a.x, s * a.y, s * a.z); }
LB_HD V3 operator/(V3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
LB_HD V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
LB_HD float dot3(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
LB_HD V3 cross3(V3 a, V3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
LB_HD float lensq3(V3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }
LB_HD float len3(V3 a) { return sqrtf(lensq3(a)); }
LB_HD V3 norm3(V3 a)
{
    float l2 = lensq3(a);
    if (l2 == 0) return mk3(0.f);
    float inv = 1.0f / sqrtf(l2);
    return mk3(a.x * inv, a.y * inv, a.z * inv);
}
LB_HD bool near_zero3(V3 a) { return fabsf(a.x) < LB_SMALL && fabsf(a.y) < LB_SMALL && fabsf(a.z) < LB_SMALL; }
LB_HD bool is_zero3(V3 a) { return a.x == 0 && a.y == 0 && a.z == 0; }
LB_HD float fminr(float a, float b) { return a < b ? a : b; }   /* TMIN: a < b ? a : b */
LB_HD float fmaxr(float a, float b) { return a > b ? a : b; }   /* TMAX: a > b ? a : b */
LB_HD V3 min3(V3 a, V3 b) { return mk3(fminr(a.x, b.x), fminr(a.y, b.y), fminr(a.z, b.z)); }
LB_HD V3 max3(V3 a, V3 b) { return mk3(fmaxr(a.x, b.x), fmaxr(a.y, b.y), fmaxr(a.z, b.z)); }
/* TLERP(a,b,s) = a*(1-s) + b*s  (lighter_int.hpp:285) */
LB_HD V3 lerp3(V3 a, V3 b, float s) { return a * (1.0f - s) + b * s; }
LB_HD float lerpf(float a, float b, float s) { return a * (1.0f - s) + b * s; }

struct Box3 { V3 lo, hi; };
LB_HD bool box_valid(const Box3 &b) { return b.lo.x <= b.hi.x && b.lo.y <= b.hi.y && b.lo.z <= b.hi.z; }

/*
 * rad_cull.h -- the conservative culling tests of the radiosity pair sweep (host + device).
 *
 * The reference tests every lumel pair (lighter.cpp:735-752: dotA = Ni.d, dotB = Nj.(-d), both > 0.001, factor
 * dotA*dotB/(len^4*pi) >= 0.001).  The sweep (gpu_radiosity.cu) skips whole blocks of pairs when interval bounds prove that
 * no pair of the block can pass; these tests must never reject a block that holds a linking pair.  They are shared with the
 * host so that tests/test_host.py can check exactly that against the oracle's pair criterion, without a GPU.
 */
#pragma once
#include <vector_types.h>
#include <vector_functions.h>    /* make_float4 on the host */
#include "vmath.h"

#ifndef LB_RAD_FLATN
#define LB_RAD_FLATN 0          /* 1 = flat-normal fast path of row_group_may_link (A/B variant) */
#endif
#ifndef LB_RAD_TWOSTAGE
#define LB_RAD_TWOSTAGE 0       /* 1 = the sweep's per-row group test votes after its cheap half (A/B variant) */
#endif
#define RAD_CUTOFF 17.85f       /* > sqrt(1/(0.001*pi)) = 17.8412; conservative */
#define RAD_SKIP_BELOW 0.0009f  /* interval bounds below this cannot reach the 0.001 thresholds even with rounding */

struct TileBounds { float4 plo, phi, nlo, nhi; };      /* 64 bytes: position box, normal-component box */

/* max over n in [nl,nh], d in [dl,dh] of n*d (one axis) */
LB_HD float imax_prod(float nl, float nh, float dl, float dh)
{
    return fmaxf(fmaxf(nl * dl, nl * dh), fmaxf(nh * dl, nh * dh));
}

/* can any lumel pair of the two tiles link?  Conservative (never rejects a linking pair). */
LB_HD bool tile_pair_may_link(const TileBounds &R, const TileBounds &C)
{
    /* d = P_c - P_r per axis */
    const float dlx = C.plo.x - R.phi.x, dhx = C.phi.x - R.plo.x;
    const float dly = C.plo.y - R.phi.y, dhy = C.phi.y - R.plo.y;
    const float dlz = C.plo.z - R.phi.z, dhz = C.phi.z - R.plo.z;
    const float gx = fmaxf(fmaxf(dlx, -dhx), 0.f), gy = fmaxf(fmaxf(dly, -dhy), 0.f), gz = fmaxf(fmaxf(dlz, -dhz), 0.f);
    const float min_len2 = gx * gx + gy * gy + gz * gz;
    if (!(min_len2 <= RAD_CUTOFF * RAD_CUTOFF)) return false;         /* also rejects padding tiles (inf/nan) */
    const float maxA = imax_prod(R.nlo.x, R.nhi.x, dlx, dhx) + imax_prod(R.nlo.y, R.nhi.y, dly, dhy) + imax_prod(R.nlo.z, R.nhi.z, dlz, dhz);
    const float maxB = imax_prod(C.nlo.x, C.nhi.x, -dhx, -dlx) + imax_prod(C.nlo.y, C.nhi.y, -dhy, -dly) + imax_prod(C.nlo.z, C.nhi.z, -dhz, -dlz);
    if (maxA < RAD_SKIP_BELOW || maxB < RAD_SKIP_BELOW) return false;
    if (min_len2 > 0.f && maxA * maxB < RAD_SKIP_BELOW * 3.14159265f * min_len2 * min_len2) return false;
    return true;
}

/* can row lumel (P, N) link with any lumel of the group bounded by C?  Conservative, like tile_pair_may_link: the maximum of
 * the row's linear form N.(c - P) over the group's position box is exact; the group's side uses interval products. */
LB_HD bool row_group_may_link(const V3 &P, const V3 &N, const TileBounds &C)
{
    const float lx = C.plo.x - P.x, hx = C.phi.x - P.x, ly = C.plo.y - P.y, hy = C.phi.y - P.y, lz = C.plo.z - P.z, hz = C.phi.z - P.z;   /* d = c - P per axis */
    const float gx = fmaxf(fmaxf(lx, -hx), 0.f), gy = fmaxf(fmaxf(ly, -hy), 0.f), gz = fmaxf(fmaxf(lz, -hz), 0.f);
    const float min_len2 = gx * gx + gy * gy + gz * gz;
    if (!(min_len2 <= RAD_CUTOFF * RAD_CUTOFF)) return false;
    const float maxA = fmaxf(N.x * lx, N.x * hx) + fmaxf(N.y * ly, N.y * hy) + fmaxf(N.z * lz, N.z * hz);
    float maxB;
#if LB_RAD_FLATN
    /* a group on a flat face has ONE normal (nlo == nhi, flagged in plo.w by rad_tile_bounds_kernel): the four interval
     * products per axis collapse to two, with the same value bit for bit.  The branch is warp-uniform (one group per step). */
    if (C.plo.w != 0.f) maxB = fmaxf(C.nlo.x * -hx, C.nlo.x * -lx) + fmaxf(C.nlo.y * -hy, C.nlo.y * -ly) + fmaxf(C.nlo.z * -hz, C.nlo.z * -lz);
    else
#endif
    maxB = imax_prod(C.nlo.x, C.nhi.x, -hx, -lx) + imax_prod(C.nlo.y, C.nhi.y, -hy, -ly) + imax_prod(C.nlo.z, C.nhi.z, -hz, -lz);
    if (maxA < RAD_SKIP_BELOW || maxB < RAD_SKIP_BELOW) return false;
    if (min_len2 > 0.f && maxA * maxB < RAD_SKIP_BELOW * 3.14159265f * min_len2 * min_len2) return false;
    return true;
}


/* row_group_may_link in two steps for the sweep (LB_RAD_TWOSTAGE): step A decides on the distance and on the row's own side
 * (maxA, exact per row -- this is what the per-row test adds over the warp-level interval test), step B adds the group's side
 * and the factor bound.  A AND B == row_group_may_link; the sweep votes after A and only evaluates B when some lane passed. */
struct RowGroupA { float lx, hx, ly, hy, lz, hz, min_len2, maxA; };
LB_HD bool row_group_step_a(const V3 &P, const V3 &N, const TileBounds &C, RowGroupA &t)
{
    t.lx = C.plo.x - P.x; t.hx = C.phi.x - P.x; t.ly = C.plo.y - P.y; t.hy = C.phi.y - P.y; t.lz = C.plo.z - P.z; t.hz = C.phi.z - P.z;
    const float gx = fmaxf(fmaxf(t.lx, -t.hx), 0.f), gy = fmaxf(fmaxf(t.ly, -t.hy), 0.f), gz = fmaxf(fmaxf(t.lz, -t.hz), 0.f);
    t.min_len2 = gx * gx + gy * gy + gz * gz;
    if (!(t.min_len2 <= RAD_CUTOFF * RAD_CUTOFF)) return false;
    t.maxA = fmaxf(N.x * t.lx, N.x * t.hx) + fmaxf(N.y * t.ly, N.y * t.hy) + fmaxf(N.z * t.lz, N.z * t.hz);
    return !(t.maxA < RAD_SKIP_BELOW);
}
LB_HD bool row_group_step_b(const TileBounds &C, const RowGroupA &t)
{
    const float maxB = imax_prod(C.nlo.x, C.nhi.x, -t.hx, -t.lx) + imax_prod(C.nlo.y, C.nhi.y, -t.hy, -t.ly) + imax_prod(C.nlo.z, C.nhi.z, -t.hz, -t.lz);
    if (maxB < RAD_SKIP_BELOW) return false;
    if (t.min_len2 > 0.f && t.maxA * maxB < RAD_SKIP_BELOW * 3.14159265f * t.min_len2 * t.min_len2) return false;
    return true;
}

/* lumel x lumel fast filter of the sweep's lock-step phase: an FMA dot differs from the reference's mul/add dot by < 1e-5 for
 * any pair close enough to link (|d| <= 17.85), and the factor inequality dr*dj >= 0.001*pi*len^4 is evaluated with relative
 * error ~1e-6: with every threshold lowered by 10 % no linking pair is lost.  Survivors are re-evaluated exactly. */
LB_HD float rad_fma(float a, float b, float c)
{
#if defined(__CUDA_ARCH__)
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
LB_HD bool rad_fast_filter(const V3 &Pr, const V3 &Nr, const float4 &pj, const float4 &nj)
{
    const float dx = pj.x - Pr.x, dy = pj.y - Pr.y, dz = pj.z - Pr.z;
    const float drf = rad_fma(Nr.z, dz, rad_fma(Nr.y, dy, Nr.x * dx));
    const float djf = -rad_fma(nj.z, dz, rad_fma(nj.y, dy, nj.x * dx));
    const float l2 = rad_fma(dz, dz, rad_fma(dy, dy, dx * dx));
    return fminf(drf, djf) > RAD_SKIP_BELOW && drf * djf >= (RAD_SKIP_BELOW * 3.14159265f) * (l2 * l2);
}

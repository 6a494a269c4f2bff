// kernels_fused_x2.cuh -- the Float32 performance path: the fused RK2 stage kernels with TWO columns per
// thread and Blackwell's packed fp32 arithmetic (FFMA2 / FADD2 / FMUL2, sm_100a only).
//
// Why: ncu shows the one-column kernels (kernels_fused.cuh) are bound by instruction ISSUE, not by HBM
// (profiles/r01b: issue-active 85 %, DRAM 29 %).  Half of the issued instructions are FADD/FMUL/FFMA, so
//   * every add/mul/fma on (column 2t, column 2t+1) is ONE packed instruction (scratch/ubench.cu measures
//     FFMA2+FMNMX mixes at 1.3-1.4 cycles/instr vs 1.3-1.5 for the scalar mix: twice the flops per slot);
//   * x-neighbours are half in the thread's own registers: 12 shuffles per 2 cells instead of per cell;
//   * a warp covers 64 columns and produces 60 (6 % overlap instead of 12.5 %); loads/stores are 64-bit;
//   * the physics is re-associated for the select-heavy parts (fast policy only, tolerance 1e-5):
//       - dry face sides are masked on the INPUT (h*m, (1/h*)*m with m = 0/1 from one FSET) instead of five
//         selects on the outputs; h >= 0 after masking, so the desingularisation needs no copysign;
//       - a- is carried as -a- = max(c-u, ., 0), which removes the negations of the central-upwind formula;
//       - the slope clamp  (w - s/2 < Bl ? 2(w-Bl) : w + s/2 < Br ? 2(Br-w) : s)  is evaluated as
//         (s > s1 ? s1 : max(s, s2)) with s1 = 2w - 2Bl, s2 = 2Br - 2w (same predicate, no face value needed);
//       - minmod(a,b,c) = min(max(min3,0),max3): 2 FMNMX3 + 2 FMNMX, no add;
//       - 2*B_face = sum of two corners is used directly (the 0.5 is folded into constants).
// Semantics are those of kernels_fused.cuh (same reference citations); parity: tests/test_gpu_parity.py
// compares this path against the CPU oracle at rel-Linf <= 1e-5 (the north_star Float32 tolerance).
#pragma once
#include <cstddef>
#include <type_traits>

#include "kernels_fused.cuh"

SWE_NS_BEGIN
namespace x2 {

typedef float2 f2;

constexpr int VW = 60;       // valid (output) columns per warp: lanes 1..30, two columns each
constexpr int WPB = 4;       // warps per block
constexpr int ROWS = 66;     // rows marched per block

__device__ __forceinline__ f2 mk(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ f2 sp(float a) { return make_float2(a, a); }
__device__ __forceinline__ f2 add(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 mul(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 sub(f2 a, f2 b) { return __ffma2_rn(b, sp(-1.0f), a); }   // a - b
__device__ __forceinline__ f2 swp(f2 a) { return make_float2(a.y, a.x); }
__device__ __forceinline__ f2 neg(f2 a) { return make_float2(-a.x, -a.y); }
// X2_ACCURATE=1: correctly rounded reciprocal / square root / libm cbrt instead of the MUFU approximations
// (<= 1-2 ulp each).  An A/B switch for the parity study (profiles/r02_parity.md): it separates the approximation
// error of these three from the conditioning of the wet/dry scheme.
#ifndef X2_ACCURATE
#define X2_ACCURATE 0
#endif
__device__ __forceinline__ float rcp_a(float x) {
#if X2_ACCURATE
    return __frcp_rn(x);
#endif
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sqrt_a(float x) {
#if X2_ACCURATE
    return __fsqrt_rn(x);
#endif
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float cbrt_a(float x) {   // sign(x) * 2^(log2|x| / 3): two MUFU
#if X2_ACCURATE
    return cbrtf(x);
#endif
    float l, e;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l) : "f"(fabsf(x)));
    l *= 0.33333334f;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(l));
    return copysignf(e, x);
}

// kernel parameters: the Float64 scalars of the reference as packed Float32 pairs (constant bank)
struct Par2 {
    int Nx, Ny;
    long long plane;
    int Bp;                       // pitch of the corner array
    int bc, friction, bottom_source, ylo_phys, yhi_phys;
    int rain_kind, infil_kind, infil_first, rows;
    int nitems;                   // work items of one launch (Args2::items)
    float cutoff, kappa, constB;
    f2 theta, hith /*0.5/theta*/, half, mhalf, quarter, g, K /*0.5 rho g*/, i2k, hk, rho, nrdx, nrdy, prdy /*+1/dy*/, botx, boty, Cz;
};

struct Args2 {
    Par2 p;
    const float* U;
    const float* D;
    float* out;
    const float* B;
    const float* wrapLo;          // corner rows of the wrap-around image's source rows (see StageArgs), or nullptr
    const float* wrapHi;
    const double* factor;
    Ctl* ctl;
    const int4* items;            // (tile, first interior row, number of rows (multiple of 3 except at the top), 0)
};

// ---- minmod(a,b,c), linearreconstruction.jl:28-35, as the median of (min3, 0, max3) ----------------
__device__ __forceinline__ float minmod3f(float a, float b, float c) {
    const float lo = fminf(fminf(a, b), c), hi = fmaxf(fmaxf(a, b), c);
    return fminf(fmaxf(lo, 0.0f), hi);
}

// ---- reconstruction of a pair of cells from the centre values and the backward / forward differences
// (linearreconstruction.jl:37-43, 73-100; noreconstruction.jl:30-46).  tBl/tBr = 2 * face bathymetry.
// faces2: slope clamp of w (no negative face depth), face values, depth = w - face bathymetry
__device__ __forceinline__ void faces2(const Par2& p, const f2 c[3], f2 s[3], f2 tBl, f2 tBr, f2 L[3], f2 R[3]) {
    const f2 w2 = add(c[0], c[0]);
    const f2 s1 = sub(w2, tBl), s2 = sub(tBr, w2);
    s[0].x = s[0].x > s1.x ? s1.x : fmaxf(s[0].x, s2.x);
    s[0].y = s[0].y > s1.y ? s1.y : fmaxf(s[0].y, s2.y);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        L[k] = fma2(s[k], p.mhalf, c[k]);
        R[k] = fma2(s[k], p.half, c[k]);
    }
    L[0] = fma2(tBl, p.mhalf, L[0]);
    R[0] = fma2(tBr, p.mhalf, R[0]);
}
// y direction: the backward / forward differences are pairs of the same two columns
template <int RECON>
__device__ __forceinline__ void recon2(const Par2& p, const f2 c[3], const f2 bwd[3], const f2 fwd[3], f2 tBl, f2 tBr,
                                       f2 Bc, f2 L[3], f2 R[3]) {
    if (RECON == 0) {
        L[0] = R[0] = sub(c[0], Bc);
        L[1] = R[1] = c[1];
        L[2] = R[2] = c[2];
        return;
    }
    f2 s[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const f2 tf = mul(p.theta, fwd[k]), tb = mul(p.theta, bwd[k]);
        const f2 cen = mul(p.half, add(fwd[k], bwd[k]));
        s[k] = mk(minmod3f(tf.x, cen.x, tb.x), minmod3f(tf.y, cen.y, tb.y));
    }
    faces2(p, c, s, tBl, tBr, L, R);
}
// X2_TCARRY=1: the y march carries theta * (q(j+1) - q(j)) instead of the raw difference, so the backward limiter
// argument of cell j+1 is the forward one of cell j (one FMUL2 per variable and row less); the central
// difference is rebuilt as (0.5/theta) * (tf + tb).
#ifndef X2_TCARRY
#define X2_TCARRY 3          // bit 0: stage 1, bit 1: stage 2 (round 2: stage 1 too, K1 -1.7 %; round 1 it cost registers there)
#endif
#ifndef X2_DRYLEAN
#define X2_DRYLEAN 0         // 1: all-dry warp-rows skip the x reconstruction of the momenta and the flux differences (see the march).
                             // Bitwise the same results and -18 % / -26 % instructions on a dry row of K2 / K1, but dry rows wait for memory, not
                             // for issue slots: K1 -1.9 %, K2 +1.8 % on the bench line, +1.6 % with every cell wet (profiles/r03b_ab_diet.txt): off
#endif
#ifndef X2_NOGHOST_FAST
#define X2_NOGHOST_FAST 0    // 1: warps holding x-ghost pairs never take the FAST trips (which then drop the ghost code): K2 -1.6 .. -2 % at
                             // 16384^2, +2 % at 8192^2, +13 % at 4096^2 where 2 of 69 tile columns hold ghosts (profiles/r03b_ab_diet.txt): off.
                             // 2: K2 carries BOTH FAST flavours and every warp takes the one it needs: bitwise the same results, K2 -2.8 %
                             // at 16384^2 but +13 % at 4096^2 for the same instantiation (unexplained; profiles/r03b_ab_diet.txt): off
#endif
template <int RECON>
__device__ __forceinline__ void recon2t(const Par2& p, const f2 c[3], const f2 tb[3], const f2 fwd[3], f2 tBl, f2 tBr,
                                        f2 Bc, f2 L[3], f2 R[3], f2 tf[3]) {
    if (RECON == 0) {
        L[0] = R[0] = sub(c[0], Bc);
        L[1] = R[1] = c[1];
        L[2] = R[2] = c[2];
        return;
    }
    f2 s[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        tf[k] = mul(p.theta, fwd[k]);
        const f2 cen = mul(p.hith, add(tf[k], tb[k]));
        s[k] = mk(minmod3f(tf[k].x, cen.x, tb[k].x), minmod3f(tf[k].y, cen.y, tb[k].y));
    }
    faces2(p, c, s, tBl, tBr, L, R);
}
// x direction: the two cells of a thread are neighbours, so the three differences d0 = a - left, d1 = b - a,
// d2 = right - b are shared; everything is built from scalars that land in aligned pairs (no pair shuffling).
// (components K0 .. K1 only: the depth (0) decides whether a row's faces are wet; the momenta (1, 2) are only needed if they are)
template <int RECON, int K0 = 0, int K1 = 2>
__device__ __forceinline__ void recon2x(const Par2& p, const f2 c[3], const float ql[3], const float qr[3], f2 tBl,
                                        f2 tBr, f2 Bc, f2 L[3], f2 R[3]) {
    if (RECON == 0) {
#pragma unroll
        for (int k = K0; k <= K1; ++k) L[k] = R[k] = k == 0 ? sub(c[0], Bc) : c[k];
        return;
    }
    f2 s[3];
#pragma unroll
    for (int k = K0; k <= K1; ++k) {
        const f2 d01 = mk(c[k].x - ql[k], c[k].y - c[k].x);               // (d0, d1)
        const float d2 = qr[k] - c[k].y;
        const f2 t01 = mul(p.theta, d01);
        const float t2 = p.theta.x * d2;
        const f2 cen = mul(p.half, mk(c[k].y - ql[k], qr[k] - c[k].x));    // ((d0+d1)/2, (d1+d2)/2)
        s[k] = mk(minmod3f(t01.y, cen.x, t01.x), minmod3f(t2, cen.y, t01.y));
    }
    if (K0 == 0 && K1 == 2) { faces2(p, c, s, tBl, tBr, L, R); return; }
    if (K0 == 0) {                                       // faces2, depth only
        const f2 w2 = add(c[0], c[0]);
        const f2 s1 = sub(w2, tBl), s2 = sub(tBr, w2);
        s[0].x = s[0].x > s1.x ? s1.x : fmaxf(s[0].x, s2.x);
        s[0].y = s[0].y > s1.y ? s1.y : fmaxf(s[0].y, s2.y);
    }
#pragma unroll
    for (int k = K0; k <= K1; ++k) {
        L[k] = fma2(s[k], p.mhalf, c[k]);
        R[k] = fma2(s[k], p.half, c[k]);
    }
    if (K0 == 0) {
        L[0] = fma2(tBl, p.mhalf, L[0]);
        R[0] = fma2(tBr, p.mhalf, R[0]);
    }
}

// desingularised depth of a NON-NEGATIVE depth pair (equation.jl:56-62): max(h, min(h^2/2k + k/2, k)) is
// h + relu(k - h)^2 / 2k; k - h <= k < 1, so the relu is one saturating add on the FMA pipe.
// (sub.sat as ONE instruction: nvcc compiles __saturatef(k - h) into FADD + FADD.SAT -- 12 of K2's ~520 instructions per thread-row)
__device__ __forceinline__ float sub_sat(float a, float b) {
    float r;
    asm("sub.sat.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}
#ifndef X2_SUBSAT
#define X2_SUBSAT 1
#endif
__device__ __forceinline__ f2 desing_pos2(const Par2& p, f2 h) {
#if X2_SUBSAT
    const f2 rl = mk(sub_sat(p.kappa, h.x), sub_sat(p.kappa, h.y));
#else
    const f2 rl = mk(__saturatef(p.kappa - h.x), __saturatef(p.kappa - h.y));
#endif
    return fma2(mul(rl, rl), p.i2k, h);
}

// ---- one side of a pair of faces: physical flux, u+c and c-u with the dry gate folded into the inputs
// (swe_2D.jl:43-81, equation.jl:40-62, centralupwind.jl:54-63) ------------------------------------------
template <int DIR>
__device__ __forceinline__ void side2(const Par2& p, f2 h, f2 hu, f2 hv, f2 F[3], f2& ep, f2& en) {
    const f2 m = mk(h.x > p.cutoff ? 1.0f : 0.0f, h.y > p.cutoff ? 1.0f : 0.0f);
    const f2 hm = mul(h, m);
    const f2 hh = mul(hm, hm);
    const f2 pres = mul(p.K, hh);
    const f2 gh = mul(p.g, hm);
    const f2 c = mk(sqrt_a(gh.x), sqrt_a(gh.y));
    const f2 hs = desing_pos2(p, hm);
    const f2 rm = mul(mk(rcp_a(hs.x), rcp_a(hs.y)), m);
    const f2 u = mul(hu, rm), v = mul(hv, rm);
    const f2 rh = mul(p.rho, hm);
    f2 un;
    if (DIR == 0) {
        F[0] = mul(rh, u);
        F[1] = fma2(F[0], u, pres);
        F[2] = mul(F[0], v);
        un = u;
    } else {
        F[0] = mul(rh, v);
        F[1] = mul(F[0], u);
        F[2] = fma2(F[0], v, pres);
        un = v;
    }
    ep = add(un, c);
    en = sub(c, un);
}

// ---- Kurganov-Petrova central-upwind flux through a pair of faces (centralupwind.jl:48-77) -------------
// (The caller skips the evaluation for a warp whose faces are all dry on both sides.)
__device__ __forceinline__ bool any_wet(const Par2& p, f2 hm, f2 hp) {
    return fmaxf(fmaxf(hm.x, hm.y), fmaxf(hp.x, hp.y)) > p.cutoff;
}
#ifndef X2_CU12
#define X2_CU12 1            // bit 0: stage 1, bit 1: stage 2 (measured: helps K1, hurts K2's register allocation)
#endif
template <int DIR, bool SPEED>
__device__ __forceinline__ void cu2(const Par2& p, const f2 Um[3], const f2 Up[3], f2 F[3], f2& speed) {
    f2 Fm[3], Fp[3], epm, enm, epp, enp;
    side2<DIR>(p, Um[0], Um[1], Um[2], Fm, epm, enm);
    side2<DIR>(p, Up[0], Up[1], Up[2], Fp, epp, enp);
    const f2 ap = mk(fmaxf(fmaxf(epm.x, epp.x), 0.0f), fmaxf(fmaxf(epm.y, epp.y), 0.0f));   //  a+
    const f2 an = mk(fmaxf(fmaxf(enm.x, enp.x), 0.0f), fmaxf(fmaxf(enm.y, enp.y), 0.0f));   // -a-
    const f2 d = add(ap, an);
    const bool okx = !(d.x < p.kappa), oky = !(d.y < p.kappa);                            // :67-69
    const f2 r = mk(okx ? rcp_a(d.x) : 0.0f, oky ? rcp_a(d.y) : 0.0f);
    const f2 wp = mul(ap, r), wn = mul(an, r);
    constexpr bool C12 = SPEED ? (X2_CU12 & 1) : ((X2_CU12 >> 1) & 1);      // SPEED <=> stage 1
    if (C12) {
        // wp (F- + (-a-)(U- - U+)) + wn F+ : the same flux with the dissipation folded into the left term (12 instead of 13)
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const f2 jn = sub(Um[k], Up[k]);
            const f2 t = mul(wp, fma2(an, jn, Fm[k]));
            F[k] = fma2(wn, Fp[k], t);
        }
    } else {
        const f2 c2 = mul(ap, wn);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const f2 jn = sub(Um[k], Up[k]);
            f2 t = mul(wp, Fm[k]);
            t = fma2(wn, Fp[k], t);
            F[k] = fma2(c2, jn, t);
        }
    }
    if (SPEED) speed = mk(okx ? fmaxf(ap.x, an.x) : 0.0f, oky ? fmaxf(ap.y, an.y) : 0.0f);
    else speed = sp(0.0f);
}

// ---- pointwise tail of a substep on a pair of cells (rungekutta.jl:39-41, friction.jl:62-84,
// timestepper.jl:41-51): returns the factor the momenta are multiplied by -- 1/(1 - dt*c) of the implicit
// friction times the wet flag of the dry clean-up -- and raises o0 to the cell bathymetry (for a wet cell
// o0 - Bc >= cutoff >= 0, so the max is the identity there).
#ifndef X2_FRIC3
#define X2_FRIC3 1
#endif
template <int FRIC>
__device__ __forceinline__ f2 tail_factor2(const Par2& p, const f2 prev[3], f2 Bc, f2 czdt, f2& o0) {
    const f2 dpt = sub(o0, Bc);
    f2 fac = mk(dpt.x < p.cutoff ? 0.0f : 1.0f, dpt.y < p.cutoff ? 0.0f : 1.0f);
    o0 = mk(fmaxf(o0.x, Bc.x), fmaxf(o0.y, Bc.y));
    if (FRIC != 0) {
        const f2 h = sub(prev[0], Bc);
        const f2 ha = mk(fabsf(h.x), fabsf(h.y));
        const f2 hsa = desing_pos2(p, ha);                       // |h*|; h* = copysign(|h*|, h)
        // speed = sqrt(u^2+v^2) = sqrt(hu^2+hv^2)/|h*|
        const f2 m2 = fma2(prev[1], prev[1], mul(prev[2], prev[2]));
        const f2 sm = mk(sqrt_a(m2.x), sqrt_a(m2.y));
#if X2_FRIC3
        if (FRIC == 1) {
            // friction_bh2021: 1/(1 + dt Cz |m|/h*^3) = h*^3 / (h*^3 + dt Cz |m|): one reciprocal instead of two
            const f2 h3 = mul(mul(hsa, hsa), hsa);
            const f2 den = fma2(sm, czdt, h3);
            return mul(fac, mul(h3, mk(rcp_a(den.x), rcp_a(den.y))));
        }
#endif
        const f2 ra = mk(rcp_a(hsa.x), rcp_a(hsa.y));            // 1/|h*|
        const f2 s = mul(sm, ra);
        f2 iden;                                                  // 1 / (law denominator), positive for every law
        if (FRIC == 1) iden = mul(ra, ra);                       // h*^2              (friction_bh2021)
        else {
            const f2 rc = mk(rcp_a(cbrt_a(hsa.x)), rcp_a(cbrt_a(hsa.y)));
            iden = mul(rc, ra);                                  // cbrt(h*) h*       (friction_bsa2012)
            if (p.friction != SWE_B200_FRICTION_BSA2012) iden = mul(iden, ra);   // cbrt(h*) h*^2 (friction_fcg2016)
        }
        const f2 den = fma2(mul(s, iden), czdt, sp(1.0f));       // 1 - dt*(-Cz*speed/denominator)
        fac = mul(fac, mk(rcp_a(den.x), rcp_a(den.y)));
    }
    return fac;
}

// ---- asynchronous global -> shared copies (LDGSTS): a prefetched row costs no registers while in flight ----
__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem)
                 : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem)
                 : "memory");
}
__device__ __forceinline__ void cp_async16_cg(void* smem, const void* gmem) {     // bypasses L1 (.cg needs 16 bytes)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem)
                 : "memory");
}
// X2_FILL16=1 (experiment, profiles/r02_fill16.md): in FAST trips of warps whose 64 columns are all in range and hold no
// remapped ghost pair, a state row is fetched as 16 x 16-byte cp.async.cg chunks per array, spread over the lanes
// (3 copies per lane and row in stage 2 instead of 6, 1.5 instead of 3 in stage 1; no L1 allocation), at the price of
// one __syncwarp per row (a lane then reads ring data another lane copied).  Needs 16-byte aligned rows (Nx % 4 == 0).
#ifndef X2_FILL16
#define X2_FILL16 0
#endif
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// output rows are never read again by the launch that writes them: X2_STCS=1 stores them evict-first (st.global.cs)
#ifndef X2_STCS
#define X2_STCS 0
#endif
__device__ __forceinline__ void st_out(char* p, f2 v) {
#if X2_STCS
    __stcs(reinterpret_cast<float2*>(p), v);
#else
    *reinterpret_cast<f2*>(p) = v;
#endif
}

constexpr int NT = WPB * 32;
#ifndef X2_MINB
#define X2_MINB 4            // resident blocks per SM, stage 1 (128 registers)
#endif
#ifndef X2_MINB2
#define X2_MINB2 3           // stage 2 carries more state: 3 blocks x 168 registers (no spills) measured 1.7 % faster than 4 x 128
#endif
template <int STAGE>
constexpr int minb() { return STAGE == 2 ? X2_MINB2 : X2_MINB; }

// Per-thread rings in shared memory.  Every thread only ever reads what it copied itself (its own two
// columns), so cp.async.wait_group is the only synchronisation: there is no barrier in the kernel.
// The march is unrolled by 3 (the period of the 3-row register window), so a ring of depth 3 has compile-time
// slots; depth-6 rings are 2 banks x 3 slots and the bank flips once per unrolled body.
//   U : U^n rows.  Stage 2 reads row j again for the Heun average while rows j+3, j+4 are in flight -> depth 6
//       (stage 1: depth 3).
//   D : dU rows (stage 2): row j+2 is consumed on arrival, rows j+3, j+4 in flight            -> depth 3.
//   E : bathymetry corner rows: j+1 (x faces), j+3 (arrived), j+4, j+5 (in flight)            -> depth 6.
template <int STAGE>
struct Rings {
    f2 U[STAGE == 2 ? 2 : 1][3][3][NT];      // [bank][slot][variable][thread]
    f2 D[STAGE == 2 ? 3 : 1][3][STAGE == 2 ? NT : 1];
    f2 E[2][3][NT];
    float E2[2][3][NT];
};

// flavour of a 3-row trip: value = FAST (every fetched row is an interior row), noghost = the warp holds no x-ghost pair either
template <bool F, bool NG>
struct TripTag { static constexpr bool value = F; static constexpr bool noghost = NG; };

template <int RECON, bool HASB, int STAGE, int FRIC, bool FACTOR>
__global__ void __launch_bounds__(NT, minb<STAGE>()) k_stage_x2(const Args2 a) {
    __shared__ Rings<STAGE> sm;
    constexpr int UB = STAGE == 2 ? 2 : 1;               // banks of the U ring
    const Par2& p = a.p;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int Nx = p.Nx, Ny = p.Ny;
    f2 msx = sp(0.0f);                                   // running CFL maxima of this thread (stage 1): x faces a / b,
    float msy = 0.0f;                                    // y faces of both columns (same predicate: one FMNMX3 per row)
    // Work queue: every WARP is independent (its own ring columns, no barrier), so warps pull (tile, rows)
    // items from one atomic counter until the grid is done.  The item table is a guided schedule (EngineT::
    // build_items): 126-row items first, then 63-, 30- and 12-row items, so that all 148 SMs stay busy to the
    // end of the launch although dry stretches finish ~40 % sooner than wet ones.
    const int total_items = p.nitems;
  for (;;) {
    int item = 0;
    if (lane == 0) item = (int)atomicAdd(&a.ctl->queue[STAGE - 1], 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= total_items) break;
    const int4 it = a.items[item];                       // (tile, first row, rows, -): guided schedule, large items first
    const int tile = it.x;
    const int i = tile * VW + 2 * lane;                  // first of this lane's two columns (even)
    const int ic = min(i, Nx - 2);                       // clamped for loads (Nx is even)
    const int ja = it.y;
    const int jb = min(ja + it.z, Ny - 2);
    const bool in_x = i <= Nx - 4;
    const bool out_pair = lane >= 1 && lane <= 30 && in_x;
    const bool facea_ok = out_pair;                      // face (i | i+1)   belongs to an interior cell
    const bool faceb_ok = lane <= 30 && in_x;            // face (i+1 | i+2)
    const long long plane = p.plane;
    const int Bp = p.Bp;
    // does this warp hold x-ghost pairs (columns 0,1 or Nx-2,Nx-1)?  warp-uniform
    const bool ghost_warp = STAGE == 2 && (tile == 0 || tile * VW + 64 > Nx - 2);
    const bool fill16 = X2_FILL16 && !ghost_warp && tile * VW + 64 <= Nx && (Nx & 3) == 0;     // warp-uniform

    f2 dt = sp(0.0f), czdt = sp(0.0f);
    int xs = ic;
    bool xswap = false;
    int xmode = 0;                                       // ghost pair: 0 mirrored + normal momentum flipped (wall), 1 mirrored (Neumann low), 2 outermost ghost twice (Neumann high)
    if (STAGE == 2) {
        const float dts = (float)a.ctl->dt_s;
        dt = sp(dts);
        czdt = mul(p.Cz, dt);
        if (ic < 2 || ic >= Nx - 2) {                    // a pair of x ghosts: where the BC copies them from
            const bool lo = ic < 2;
            if (p.bc == SWE_B200_BC_PERIODIC) xs = lo ? Nx - 4 : 2;
            else if (p.bc == SWE_B200_BC_NEUMANN) {        // bc.jl:51-53: low = mirrored pair (no sign flip); high = both
                xswap = true;                              // ghosts hold the self-evolving outermost ghost column
                xmode = lo ? 1 : 2;                        // (k_neumann_dU / k_neumann_ghosts): its own pair, .y twice
                if (lo) xs = 2;
            } else { xs = lo ? 2 : Nx - 4; xswap = true; }   // wall: mirrored pair, normal momentum flips
        }
    }
    // per-cell sources (system.jl:71-73): rain - infiltration is one constant per step unless a Horton factor plane exists
    const float rain = p.rain_kind != SWE_B200_RAIN_NONE ? (float)a.ctl->rain : 0.0f;
    const float infil_t = p.infil_kind != SWE_B200_INFIL_NONE ? (float)a.ctl->infil : 0.0f;
    const f2 src = sp(FACTOR ? rain : rain - infil_t);
    const f2 ninfil = sp(-infil_t);

    struct Corner { f2 e; float e2; };                   // B[ic..ic+2, row]
    auto ty_of = [&](const Corner& c) -> f2 {            // 2 * B_face_y of the two cells above that corner row
        return HASB ? mk(c.e.x + c.e.y, c.e.y + c.e2) : sp(2.0f * p.constB);
    };
    auto bcell = [&](f2 tylo, f2 tyhi) -> f2 { return HASB ? mul(p.quarter, add(tylo, tyhi)) : sp(p.constB); };
    auto row_source = [&](int j, bool& yflip) -> int {   // stage 2: the row the y BC copies row j from
        yflip = false;
        if (STAGE == 1) return j;
        return bc_source(j, Ny, 2, p.bc, p.ylo_phys, p.yhi_phys, yflip, 1);
    };
    // stage 1: q = U^n ; stage 2: q = U^1 = BC(postproc(friction(U^n + dt*dU))) rebuilt on the fly
    const bool xghost = xs != ic;
    // (a FAST trip -- interior rows of a warp without x-ghost pairs -- has nothing to remap, mirror or flip)
    auto finish_row = [&](auto fast_tag, bool remapped, int ys, bool yflip, const f2 u[3], const f2 d[3], f2 Bc_here, f2 q[3]) {
        constexpr bool FASTROW = decltype(fast_tag)::noghost;
        if (STAGE == 1) {
#pragma unroll
            for (int k = 0; k < 3; ++k) q[k] = u[k];
        } else {
            f2 Bm = Bc_here;
            // (halo rows only ever pass through the prologue / the generic trips: nothing is added to a FAST trip)
            const float* wr = (decltype(fast_tag)::value || !HASB) ? nullptr : (ys < 2 ? a.wrapLo : (ys >= Ny - 2 ? a.wrapHi : nullptr));
            if (!FASTROW && HASB && (xghost || remapped || wr)) {    // ghost: bathymetry of the source cells
                const float* r0 = wr ? wr + (size_t)(ys < 2 ? ys : ys - (Ny - 2)) * Bp + xs : a.B + (size_t)ys * Bp + xs;
                const float m = r0[1] + r0[Bp + 1];
                Bm = mk(0.25f * (r0[0] + r0[Bp] + m), 0.25f * (m + r0[2] + r0[Bp + 2]));
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) q[k] = fma2(dt, d[k], u[k]);
            const f2 fac = tail_factor2<FRIC>(p, u, Bm, czdt, q[0]);
            q[1] = mul(q[1], fac);
            q[2] = mul(q[2], fac);
            if (!FASTROW && ghost_warp && xswap) {
#pragma unroll
                for (int k = 0; k < 3; ++k) q[k] = xmode == 2 ? sp(q[k].y) : swp(q[k]);
                if (xmode == 0) q[1] = neg(q[1]);
            }
            if (!FASTROW && yflip) q[2] = neg(q[2]);
        }
    };

    // ---- fetch of state row r / corner row r into the rings (the first two rows of the march are requested
    // before the prologue's own loads are consumed, so the item start exposes one memory latency, not two) ----
    // Running byte pointers advance by one row per fetch.  A trip (3 rows) is FAST when every row it fetches is
    // an interior row (no y-BC remap, no clamping, every row stored); only the last trip(s) of the last chunk
    // take the generic path.
    const long long rowB = (long long)Nx * 4, planeB = plane * 4;
    const long long dOff = (const char*)a.D - (const char*)a.U;
    const char* pf = (const char*)(a.U + (size_t)(ja + 2) * Nx + xs);     // next state row to fetch
    const char* pc = (const char*)(a.B + (size_t)(ja + 3) * Bp + ic);     // next corner row to fetch
    // two rows are in flight at any time (fetch distance 2): per-row BC bookkeeping of the generic path is a 2-deep queue
    char* ps = (char*)(a.out + (size_t)ja * Nx + i);                      // row j of the output
    const double* pfac = a.factor + (size_t)ja * Nx + ic;                 // Horton factor plane, row j
    int ys_next = 0, ys_q = 0;                            // BC source row of the state rows in flight (generic path)
    bool yflip_next = false, remap_next = false, yflip_q = false, remap_q = false;
    // ring addressing in bytes: this thread's 8-byte (4-byte for E2) column inside every [.. ][NT] row
    char* const smb = reinterpret_cast<char*>(&sm);
    constexpr unsigned U_BANK = 3 * 3 * NT * 8, U_SLOT = 3 * NT * 8, U_VAR = NT * 8;
    constexpr unsigned D_OFF = offsetof(Rings<STAGE>, D), E_OFF = offsetof(Rings<STAGE>, E), E2_OFF = offsetof(Rings<STAGE>, E2);
    constexpr unsigned E_BANK = 3 * NT * 8, E_SLOT = NT * 8, E2_BANK = 3 * NT * 4, E2_SLOT = NT * 4;
    char* const smU = smb + tid * 8;
    char* const smD = smb + D_OFF + tid * 8;
    char* const smE = smb + E_OFF + tid * 8;
    char* const smE2 = smb + E2_OFF + tid * 4;
    auto fetch_row = [&](auto fast_tag, int row, unsigned ubank, int slot) {
        constexpr bool FAST = decltype(fast_tag)::value;
        const char* g = pf;
        ys_q = ys_next; yflip_q = yflip_next; remap_q = remap_next;      // the row fetched one call ago
        remap_next = false;
        yflip_next = false;
        if (!FAST) {
            ys_next = row;
            if (row >= Ny - 2) {                          // uniform
                const int rr = min(row, Ny - 1);
                ys_next = row_source(rr, yflip_next);
                remap_next = ys_next != row;
                g = (const char*)(a.U + (size_t)ys_next * Nx + xs);
            }
        }
        pf += rowB;
        if (X2_FILL16 && FAST && fill16) {
            // chunk c = lane + 32 m: array c / 16 (U planes 0..2, then dU planes), 16-byte part c % 16 of the warp's 256-byte row
            const char* gw = g - lane * 8;
            char* const sw = smb + (tid - lane) * 8;
            constexpr int NCH = (STAGE == 2 ? 6 : 3) * 16;
#pragma unroll
            for (int m = 0; m < (NCH + 31) / 32; ++m) {
                const int c = lane + 32 * m;
                if (NCH % 32 == 0 || c < NCH) {
                    const int arr = c >> 4, part = c & 15;
                    const bool isD = STAGE == 2 && arr >= 3;
                    const int k = isD ? arr - 3 : arr;
                    const char* src = gw + (isD ? dOff : 0) + k * planeB + part * 16;
                    char* dst = sw + (isD ? D_OFF : ubank) + slot * U_SLOT + k * U_VAR + part * 16;
                    cp_async16_cg(dst, src);
                }
            }
            return;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            cp_async8(smU + ubank + slot * U_SLOT + k * U_VAR, g + k * planeB);
            if (STAGE == 2) cp_async8(smD + slot * U_SLOT + k * U_VAR, g + dOff + k * planeB);
        }
    };
    auto fetch_corner = [&](auto fast_tag, int row, unsigned ebank, unsigned e2bank, int slot) {
        constexpr bool FAST = decltype(fast_tag)::value;
        if (HASB) {
            const char* g = pc;
            if (!FAST && row > Ny) g = (const char*)(a.B + (size_t)Ny * Bp + ic);
            pc += (long long)Bp * 4;
            cp_async8(smE + ebank + slot * E_SLOT, g);
            cp_async4(smE2 + e2bank + slot * E2_SLOT, g + 8);
        }
    };
    fetch_row(TripTag<false, false>{}, ja + 2, 0u, 2);
    fetch_corner(TripTag<false, false>{}, ja + 3, E_BANK, E2_BANK, 0);
    cp_async_commit();
    fetch_row(TripTag<false, false>{}, ja + 3, UB == 2 ? U_BANK : 0u, 0);
    fetch_corner(TripTag<false, false>{}, ja + 4, E_BANK, E2_BANK, 1);
    cp_async_commit();

    // ---- prologue: state rows ja-2 .. ja+1 and corner rows ja-2 .. ja+2 by direct loads ----
    f2 q[3][3];                                           // q[(j - ja) % 3] = row j of the (rebuilt) state
    f2 dprev[3], Ry0[3], Gl[3], hsy;                      // y state of cell j: q(j+1)-q(j), upper face values, lower flux, hL+hR
    Corner Ej;                                            // corner row j
    f2 tyj, ty2;                                          // 2*B_face_y above corner rows j and j+2
    {
        Corner c[5];
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            c[r].e = sp(0.0f); c[r].e2 = 0.0f;
            if (HASB) {
                const float* g = a.B + (size_t)(ja - 2 + r) * Bp + ic;
                c[r].e = *reinterpret_cast<const f2*>(g);
                c[r].e2 = g[2];
            }
        }
        f2 u[4][3], d[4][3];
        int ys[4];
        bool yf[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            ys[r] = row_source(ja - 2 + r, yf[r]);
            const size_t off = (size_t)ys[r] * Nx + xs;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                u[r][k] = *reinterpret_cast<const f2*>(a.U + off + k * plane);
                d[r][k] = STAGE == 2 ? *reinterpret_cast<const f2*>(a.D + off + k * plane) : sp(0.0f);
            }
        }
        if (STAGE == 2) {                                 // rows ja, ja+1 of U^n are read again by the Heun average
#pragma unroll
            for (int k = 0; k < 3; ++k) { sm.U[0][0][k][tid] = u[2][k]; sm.U[0][1][k][tid] = u[3][k]; }
        }
        if (HASB) {                                       // corner rows ja+1, ja+2 are read again for the x faces
            sm.E[0][1][tid] = c[3].e; sm.E2[0][1][tid] = c[3].e2;
            sm.E[0][2][tid] = c[4].e; sm.E2[0][2][tid] = c[4].e2;
        }
        const f2 tm2 = ty_of(c[0]), tm1 = ty_of(c[1]), t0 = ty_of(c[2]), t1 = ty_of(c[3]), t2 = ty_of(c[4]);
        f2 qa[3], qb[3];
        finish_row(TripTag<false, false>{}, ys[0] != ja - 2, ys[0], yf[0], u[0], d[0], bcell(tm2, tm1), qa);
        finish_row(TripTag<false, false>{}, ys[1] != ja - 1, ys[1], yf[1], u[1], d[1], bcell(tm1, t0), qb);
        finish_row(TripTag<false, false>{}, ys[2] != ja, ys[2], yf[2], u[2], d[2], bcell(t0, t1), q[0]);
        finish_row(TripTag<false, false>{}, ys[3] != ja + 1, ys[3], yf[3], u[3], d[3], bcell(t1, t2), q[1]);
        f2 b0[3], b1[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            b0[k] = sub(qb[k], qa[k]); b1[k] = sub(q[0][k], qb[k]); dprev[k] = sub(q[1][k], q[0][k]);
        }
        f2 Lt[3], Rp[3], Ly0[3];
        constexpr bool TC = (X2_TCARRY >> (STAGE - 1)) & 1;
        if (TC) {
            f2 tb0[3], tb1[3], tfd[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) tb0[k] = mul(p.theta, b0[k]);
            recon2t<RECON>(p, qb, tb0, b1, tm1, t0, bcell(tm1, t0), Lt, Rp, tb1);          // cell ja-1: upper face only
            recon2t<RECON>(p, q[0], tb1, dprev, t0, t1, bcell(t0, t1), Ly0, Ry0, tfd);     // cell ja
#pragma unroll
            for (int k = 0; k < 3; ++k) dprev[k] = RECON ? tfd[k] : dprev[k];              // carried: theta * difference
        } else {
            recon2<RECON>(p, qb, b0, b1, tm1, t0, bcell(tm1, t0), Lt, Rp);              // cell ja-1: upper face only
            recon2<RECON>(p, q[0], b1, dprev, t0, t1, bcell(t0, t1), Ly0, Ry0);         // cell ja
        }
        f2 s = sp(0.0f);
        Gl[0] = Gl[1] = Gl[2] = sp(0.0f);
        if (__any_sync(0xffffffffu, any_wet(p, Rp[0], Ly0[0]))) cu2<1, STAGE == 1>(p, Rp, Ly0, Gl, s);   // face ja-1/2
#pragma unroll
        for (int k = 0; k < 3; ++k) Gl[k] = mul(p.prdy, Gl[k]);          // carried scaled: Gl = (1/dy) * G(lower face)
        if (STAGE == 1 && out_pair) msy = fmaxf(msy, fmaxf(s.x, s.y));
        hsy = add(Ly0[0], Ry0[0]);
        Ej = c[2];
        tyj = t0;
        ty2 = t2;
    }

    // ---- march: 3 rows per trip; rows beyond jb (last chunk only) are computed on clamped data and not stored ----
    int bank = 0;                                         // (j0 - ja) / 3 mod 2
    auto trip = [&](auto fast_tag, const int j0) {
        constexpr bool FAST = decltype(fast_tag)::value;
        // byte offsets of the bank holding rows j0.. (index 0) and of the other bank (index 1)
        const unsigned uU[2] = {UB == 2 && bank ? U_BANK : 0u, UB == 2 && !bank ? U_BANK : 0u};
        const unsigned uE[2] = {bank ? E_BANK : 0u, bank ? 0u : E_BANK};
        const unsigned uE2[2] = {bank ? E2_BANK : 0u, bank ? 0u : E2_BANK};
#pragma unroll
        for (int R = 0; R < 3; ++R) {
            const int j = j0 + R;
            f2* q0 = q[R];
            f2* q1 = q[(R + 1) % 3];
            f2* q2 = q[(R + 2) % 3];
            // ---- arrivals: state row j+2 and corner row j+3 (rows j+3 / j+4 stay in flight) ----
            cp_async_wait_1();
            if (X2_FILL16) __syncwarp();                  // ring data copied by other lanes of the warp becomes visible
            // (a FAST trip does not maintain the source rows: the rows it fetches are never remapped)
            const bool yflip2 = !FAST && yflip_q, remap2 = !FAST && remap_q;
            const int ys2 = remap2 ? ys_q : j + 2;
            f2 ru[3], rd[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                ru[k] = *reinterpret_cast<const f2*>(smU + uU[((R + 2) / 3) & 1] + ((R + 2) % 3) * U_SLOT + k * U_VAR);
                rd[k] = STAGE == 2 ? *reinterpret_cast<const f2*>(smD + ((R + 2) % 3) * U_SLOT + k * U_VAR) : sp(0.0f);
            }
            Corner E1, E3;                                // corner rows j+1 and j+3
            E1.e = sp(0.0f); E1.e2 = 0.0f; E3 = E1;
            if (HASB) {
                E1.e = *reinterpret_cast<const f2*>(smE + uE[((R + 1) / 3) & 1] + ((R + 1) % 3) * E_SLOT);
                E1.e2 = *reinterpret_cast<const float*>(smE2 + uE2[((R + 1) / 3) & 1] + ((R + 1) % 3) * E2_SLOT);
                E3.e = *reinterpret_cast<const f2*>(smE + uE[1] + R * E_SLOT);
                E3.e2 = *reinterpret_cast<const float*>(smE2 + uE2[1] + R * E2_SLOT);
            }
            fetch_row(fast_tag, j + 4, uU[((R + 4) / 3) & 1], (R + 1) % 3);    // two rows ahead of the arrivals
            fetch_corner(fast_tag, j + 5, uE[((R + 5) / 3) & 1], uE2[((R + 5) / 3) & 1], (R + 2) % 3);
            cp_async_commit();
            const f2 ty1 = ty_of(E1), ty3 = ty_of(E3);
            finish_row(fast_tag, remap2, ys2, yflip2, ru, rd, bcell(ty2, ty3), q2);

            // ================= reconstruction: row j in x, cell j+1 in y =================
            // X2_DRYLEAN: the x reconstruction of the momenta (only the x flux reads it) moves into the wet branch, and an all-dry
            // warp-row finishes without the flux differences (dU = the carried lower y flux: what the zero fluxes add up to)
            f2 Lx[3], Rx[3], Ln[3], tBl, tBr, fwd[3], Ly1[3], Ry1[3];
            constexpr bool TC = (X2_TCARRY >> (STAGE - 1)) & 1;
            constexpr int KX = X2_DRYLEAN ? 0 : 2;        // last component reconstructed in x before the wet / dry decision
            f2 tfw[3];
            float ql[3], qr[3];
            {
#pragma unroll
                for (int k = 0; k <= KX; ++k) {
                    ql[k] = __shfl_up_sync(0xffffffffu, q0[k].y, 1);      // column i-1
                    qr[k] = __shfl_down_sync(0xffffffffu, q0[k].x, 1);    // column i+2
                }
                if (HASB) {
                    tBl = add(Ej.e, E1.e);
                    tBr = mk(tBl.y, Ej.e2 + E1.e2);
                } else tBl = tBr = sp(2.0f * p.constB);
                recon2x<RECON, 0, KX>(p, q0, ql, qr, tBl, tBr, bcell(tyj, ty1), Lx, Rx);
#pragma unroll
                for (int k = 0; k <= KX; ++k) Ln[k] = mk(Lx[k].y, __shfl_down_sync(0xffffffffu, Lx[k].x, 1));
#pragma unroll
                for (int k = 0; k < 3; ++k) fwd[k] = sub(q2[k], q1[k]);
                if (TC) recon2t<RECON>(p, q1, dprev, fwd, ty1, ty2, bcell(ty1, ty2), Ly1, Ry1, tfw);   // cell j+1
                else recon2<RECON>(p, q1, dprev, fwd, ty1, ty2, bcell(ty1, ty2), Ly1, Ry1);
            }
            // ================= flux differences, bottom-slope source, per-cell sources, carries, store =================
            auto row_tail = [&](auto dry_tag, const f2 F[3], const f2 G[3]) {
                constexpr bool DRY = decltype(dry_tag)::value;
                f2 dU[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    if (DRY) { dU[k] = Gl[k]; continue; }
                    const float Fp = __shfl_up_sync(0xffffffffu, F[k].y, 1);          // face (i-1 | i)
                    const f2 dF = mk(F[k].x - Fp, F[k].y - F[k].x);
                    // -(1/dy)(G - G_lower) - (1/dx) dF with G_lower carried as (1/dy) G: the carried value dies in the inner fma,
                    // so the new one (end of the row) needs no register copy
                    dU[k] = fma2(p.nrdy, G[k], fma2(p.nrdx, dF, Gl[k]));
                }
                if (HASB) {
                    dU[1] = fma2(mul(p.botx, sub(tBr, tBl)), add(Lx[0], Rx[0]), dU[1]);
                    dU[2] = fma2(mul(p.boty, sub(ty1, tyj)), hsy, dU[2]);
                }
                dU[0] = add(dU[0], src);
                if (FACTOR) {
                    const double2 fc = *reinterpret_cast<const double2*>(FAST || j < Ny ? pfac : a.factor);
                    pfac += Nx;
                    dU[0] = fma2(mk((float)fc.x, (float)fc.y), ninfil, dU[0]);
                }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    dprev[k] = (TC && RECON) ? tfw[k] : fwd[k];
                    Ry0[k] = Ry1[k]; Gl[k] = DRY ? sp(0.0f) : mul(p.prdy, G[k]);
                }
                hsy = add(Ly1[0], Ry1[0]);
                // ---- store ----
                const bool st = out_pair && (FAST || j < jb);
                if (STAGE == 1) {
                    if (st) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) st_out(ps + k * planeB, dU[k]);
                    }
                } else {
                    // rungekutta.jl:39-49 stage 2, then friction (prev = U^1) and the dry clean-up
                    f2 o[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const f2 un = *reinterpret_cast<const f2*>(smU + uU[0] + R * U_SLOT + k * U_VAR);
                        o[k] = mul(p.half, add(un, fma2(dt, dU[k], q0[k])));
                    }
                    const f2 fac = tail_factor2<FRIC>(p, q0, bcell(tyj, ty1), czdt, o[0]);
                    o[1] = mul(o[1], fac);
                    o[2] = mul(o[2], fac);
                    if (st) {
#pragma unroll
                        for (int k = 0; k < 3; ++k) st_out(ps + k * planeB, o[k]);
                    }
                }
            };
            // ================= fluxes: faces (i|i+1), (i+1|i+2) of row j and face j+1/2 =================
            // one warp-uniform decision per row: a warp whose x and y faces are all dry skips both evaluations
            const bool wet = any_wet(p, Rx[0], Ln[0]) || any_wet(p, Ry0[0], Ly1[0]);
            const bool any_w = __any_sync(0xffffffffu, wet);
            if (X2_DRYLEAN && !any_w) row_tail(std::true_type{}, nullptr, nullptr);
            else {
                f2 F[3], G[3];
                if (any_w) {
                    if (X2_DRYLEAN) {
#pragma unroll
                        for (int k = 1; k < 3; ++k) {
                            ql[k] = __shfl_up_sync(0xffffffffu, q0[k].y, 1);
                            qr[k] = __shfl_down_sync(0xffffffffu, q0[k].x, 1);
                        }
                        recon2x<RECON, 1, 2>(p, q0, ql, qr, tBl, tBr, bcell(tyj, ty1), Lx, Rx);
#pragma unroll
                        for (int k = 1; k < 3; ++k) Ln[k] = mk(Lx[k].y, __shfl_down_sync(0xffffffffu, Lx[k].x, 1));
                    }
                    f2 sx, sy;
                    cu2<0, STAGE == 1>(p, Rx, Ln, F, sx);
                    cu2<1, STAGE == 1>(p, Ry0, Ly1, G, sy);
                    if (STAGE == 1 && (FAST || j < jb)) {
                        if (facea_ok) msx.x = fmaxf(msx.x, sx.x);
                        if (faceb_ok) msx.y = fmaxf(msx.y, sx.y);
                        if (out_pair) msy = fmaxf(msy, fmaxf(sy.x, sy.y));
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 3; ++k) F[k] = G[k] = sp(0.0f);
                }
                row_tail(std::false_type{}, F, G);
            }
            ps += rowB;
            // corner state of the next row
            Ej = E1;
            tyj = ty1;
            ty2 = ty3;
        }
    };
    for (int j0 = ja; j0 < jb; j0 += 3, bank ^= 1) {
        if (j0 + 6 < Ny - 2 && !(X2_NOGHOST_FAST == 1 && ghost_warp)) {
            if (STAGE == 2 && X2_NOGHOST_FAST == 2) {     // two FAST flavours: with the ghost remap / mirror code for the warps that need it
                if (ghost_warp) trip(TripTag<true, false>{}, j0);
                else trip(TripTag<true, true>{}, j0);
            } else trip(TripTag<true, X2_NOGHOST_FAST == 1>{}, j0);
        } else trip(TripTag<false, false>{}, j0);
    }
    cp_async_wait_all();                                  // the two rows fetched past the chunk: drain before the rings are reused
  }   // work queue

    if (STAGE == 1) {
        // CFL reduction: warp shuffle -> one atomicMax per warp and direction (non-negative bit patterns)
        unsigned bx = __float_as_uint(fmaxf(msx.x, msx.y)), by = __float_as_uint(msy);
        bx = __reduce_max_sync(0xffffffffu, bx);
        by = __reduce_max_sync(0xffffffffu, by);
        if (lane == 0) {
            if (bx) atomicMax(&a.ctl->maxspeed_bits[0], (unsigned long long)bx);
            if (by) atomicMax(&a.ctl->maxspeed_bits[1], (unsigned long long)by);
        }
    }
}

}  // namespace x2
SWE_NS_END  // namespace swe

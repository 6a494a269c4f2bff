// physics.cuh -- per-cell / per-face device functions of the shallow-water step.
//
// Arithmetic policies (template parameter P):
//   Policy<float , float , false>  "fast f32": native fp32, FMA contraction, fast div/sqrt
//   Policy<double, double, false>  "fast f64": native fp64 with FMA contraction
//   Policy<float , double, true >  "strict f32": the reference's mixed Float32/Float64 promotion
//                                   (state in f32, Float64 scalars promote every expression,
//                                   rounding to f32 where the reference stores)
//   Policy<double, double, true >  "strict f64": literal IEEE sequence, no contraction
// Strict instantiations live in engine_strict.cu (compiled -fmad=false); they are written with
// operands of type S (storage) and C (compute) exactly where the reference has realtype arrays
// and Float64 scalars, so C++ promotion reproduces Julia's.
//
// Reference (paths relative to the reference's src/):
//   equations/equation.jl:40-67, equations/swe_2D.jl:43-81, equations/swe_2D_pure.jl:29-63,
//   reconstruction/linearreconstruction.jl:28-43,58-102, numericalflux/swe/centralupwind.jl:32-77,
//   friction.jl:30-78, timestepper/timestepper.jl:41-51, bottom_topography.jl:87-101.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>

#include "tu_namespace.h"

SWE_NS_BEGIN

template <class S_, class C_, bool STRICT_>
struct Policy {
    using S = S_;
    using C = C_;
    static constexpr bool strict = STRICT_;
};

// Kernel parameters: every scalar that is Float64 in the reference, in the compute type C.
template <class C>
struct Par {
    int Nx, Ny, gc;         // total cells of this strip
    long long plane;        // Nx*Ny
    C dx, dy, rdx, rdy;     // rdx = 1/dx (numericalflux.jl:27 uses 1/dx * (...))
    C theta, rho, g, half_rho_g, cutoff, kappa, two_kappa, inv_two_kappa, half_kappa, Cz, constB;
    int bc, eq_kind, friction, hasB, bottom_source;
    int ylo_phys, yhi_phys; // is the low / high y boundary of this strip a physical boundary?
};

// ---- small math helpers (overloads keep the float versions in float) ----------------------
__device__ __forceinline__ float mn(float a, float b) { return fminf(a, b); }
// Float64 has no native min/max instruction on sm_100a: fmin/fmax (IEEE NaN and signed-zero semantics) cost ~9
// instructions each, a compare-and-select 3.  The fast translation unit (SWE_FAST_MINMAX, engine_fast.cu) uses the
// latter -- identical values for finite, non-zero-tie arguments; the strict one keeps fmin/fmax (Julia's min/max).
#ifdef SWE_FAST_MINMAX
__device__ __forceinline__ double mn(double a, double b) { return a < b ? a : b; }
__device__ __forceinline__ double mx(double a, double b) { return a > b ? a : b; }
#else
__device__ __forceinline__ double mn(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ double mx(double a, double b) { return fmax(a, b); }
#endif
__device__ __forceinline__ float mx(float a, float b) { return fmaxf(a, b); }
// reciprocal for the fast policies.  Float32: a plain division (-use_fast_math makes it MUFU.RCP).  Float64: the IEEE
// division is ~20 instructions plus a slow-path call; SWE_FAST_DIV64 (engine_fast.cu) replaces it by the hardware
// seed (rcp.approx.ftz.f64, ~2^-20) and two Newton steps (4 DFMA): relative error ~1 ulp for normal arguments --
// every argument on this path is a desingularised depth, a wave-speed gap >= kappa or a friction denominator >= 1.
__device__ __forceinline__ float rcp_fast(float x) { return 1.0f / x; }
__device__ __forceinline__ double rcp_fast(double x) {
#ifdef SWE_FAST_DIV64
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
#else
    return 1.0 / x;
#endif
}
__device__ __forceinline__ float ab(float a) { return fabsf(a); }
__device__ __forceinline__ double ab(double a) { return fabs(a); }
__device__ __forceinline__ float sq(float a) { return sqrtf(a); }
__device__ __forceinline__ double sq(double a) { return sqrt(a); }
__device__ __forceinline__ float cb(float a) { return cbrtf(a); }
__device__ __forceinline__ double cb(double a) { return cbrt(a); }
__device__ __forceinline__ float cps(float a, float s) { return copysignf(a, s); }
__device__ __forceinline__ double cps(double a, double s) { return copysign(a, s); }

// ---- bathymetry lookups, 0-based (i,j) (bottom_topography.jl:87-101) ---------------------------
template <class P>
struct Bathy {
    using S = typename P::S;
    using C = typename P::C;
    const S* B;     // corner values, pitch Nx+1 (nullptr => ConstantBottomTopography)
    int pitch;
    C constB;
    __device__ __forceinline__ C cell(int i, int j) const {
        if (!B) return constB;
        const S* r0 = B + (size_t)j * pitch + i;
        const S* r1 = r0 + pitch;
        return (C)0.25 * (r0[0] + r0[1] + r1[0] + r1[1]);
    }
    __device__ __forceinline__ C face_left(int i, int j, int dir) const {
        if (!B) return constB;
        const S* r0 = B + (size_t)j * pitch + i;
        return dir == 0 ? (C)0.5 * (r0[0] + r0[pitch]) : (C)0.5 * (r0[0] + r0[1]);
    }
    __device__ __forceinline__ C face_right(int i, int j, int dir) const {
        if (!B) return constB;
        const S* r0 = B + (size_t)j * pitch + i;
        return dir == 0 ? (C)0.5 * (r0[1] + r0[pitch + 1]) : (C)0.5 * (r0[pitch] + r0[pitch + 1]);
    }
};

// ---- desingularisation, equation.jl:40-62 (Eq. 26 of Brodtkorb & Holm 2021) ---------------------
// h* = copysign(1,h) * max(|h|, min(h^2/(2k) + k/2, k));  H is S for face values, C for friction.
template <class P, class H>
__device__ __forceinline__ typename P::C desing_practical(const Par<typename P::C>& p, H h) {
    using C = typename P::C;
    H h2 = h * h;
    C inner;
    if constexpr (P::strict) inner = mn(h2 / p.two_kappa + p.half_kappa, p.kappa);
    else inner = mn((C)h2 * p.inv_two_kappa + p.half_kappa, p.kappa);
    C mag = mx((C)ab(h), inner);
    return cps(mag, (C)h);
}

// ---- physical flux and the two extreme eigenvalues in direction DIR -----------------------------
// swe_2D.jl:43-65 (flux), :69-81 (eigenvalues); Pure: swe_2D_pure.jl:29-63.
// The tangential flux is evaluated as ((rho*h)*u)*v in both directions, like the reference.
template <class P, bool PURE, int DIR>
__device__ __forceinline__ void flux_eig(const Par<typename P::C>& p, typename P::S h, typename P::S hu,
                                         typename P::S hv, typename P::C F[3], typename P::C& e1,
                                         typename P::C& e2) {
    using S = typename P::S;
    using C = typename P::C;
    S hh = h * h;
    C pres = p.half_rho_g * hh;
    C c = sq(p.g * h);
    if constexpr (PURE) {
        if (DIR == 0) {
            F[0] = p.rho * hu;
            F[1] = p.rho * hu * hu / h + pres;
            F[2] = p.rho * hu * hv / h;
        } else {
            F[0] = p.rho * hv;
            F[1] = p.rho * hu * hv / h;
            F[2] = p.rho * hv * hv / h + pres;
        }
        S un = (DIR == 0 ? hu : hv) / h;
        e1 = un + c;
        e2 = un - c;
    } else {
        C hstar = desing_practical<P>(p, h);
        C u, v;
        if constexpr (P::strict) {
            u = hu / hstar;
            v = hv / hstar;
        } else {
            C r = rcp_fast(hstar);
            u = hu * r;
            v = hv * r;
        }
        C rh = p.rho * h;
        if (DIR == 0) {
            F[0] = rh * u;
            F[1] = rh * u * u + pres;
            F[2] = rh * u * v;
        } else {
            F[0] = rh * v;
            F[1] = rh * u * v;
            F[2] = rh * v * v + pres;
        }
        C un = DIR == 0 ? u : v;
        e1 = un + c;
        e2 = un - c;
    }
}

// ---- Kurganov-Petrova central-upwind face flux, centralupwind.jl:32-45 (generic, used for the
// Pure equations) and :48-77 (practical SWE with dry-state gating) ---------------------------------
// Branch-free inside: both sides are evaluated and a dry side is masked to zero with selects (the
// reference's `if h > cutoff` gating, same values).  VOTE = true (fused kernels, all 32 lanes
// converged): a warp whose faces are all dry skips the evaluation altogether.
template <class P, bool PURE, int DIR, bool VOTE = false>
__device__ __forceinline__ void central_upwind(const Par<typename P::C>& p, const typename P::S fm[3],
                                               const typename P::S fp[3], typename P::C F[3],
                                               typename P::C& speed) {
    using S = typename P::S;
    using C = typename P::C;
    const bool wm = PURE || fm[0] > p.cutoff;
    const bool wp = PURE || fp[0] > p.cutoff;
    if (!PURE) {
        bool any = wm || wp;
        if (VOTE) any = __any_sync(0xffffffffu, any);
        if (!any) {
            F[0] = F[1] = F[2] = 0;
            speed = 0;
            return;
        }
    }
    C Fm[3], Fp[3], em1, em2, ep1, ep2;
    flux_eig<P, PURE, DIR>(p, fm[0], fm[1], fm[2], Fm, em1, em2);
    flux_eig<P, PURE, DIR>(p, fp[0], fp[1], fp[2], Fp, ep1, ep2);
    if (!PURE) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            Fm[k] = wm ? Fm[k] : (C)0;
            Fp[k] = wp ? Fp[k] : (C)0;
        }
        em1 = wm ? em1 : (C)0;
        em2 = wm ? em2 : (C)0;
        ep1 = wp ? ep1 : (C)0;
        ep2 = wp ? ep2 : (C)0;
    }
    const C ap = mx(mx(ep1, em1), (C)0);
    const C am = mn(mn(ep2, em2), (C)0);
    const C d = ap - am;
    const bool ok = PURE || !(ab(d) < p.kappa);          // centralupwind.jl:67-69
    if constexpr (P::strict) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            S jump = fp[k] - fm[k];
            const C val = (ap * Fm[k] - am * Fp[k]) / d + ((ap * am) / d) * jump;
            F[k] = ok ? val : (C)0;
        }
    } else {
        const C r = ok ? rcp_fast(d) : (C)0;
        const C c2 = ap * am * r;
#pragma unroll
        for (int k = 0; k < 3; ++k) F[k] = (ap * Fm[k] - am * Fp[k]) * r + c2 * (C)(fp[k] - fm[k]);
    }
    const bool both_dry = !PURE && fm[0] < p.cutoff && fp[0] < p.cutoff;   // :73-75
    speed = (!ok || both_dry) ? (C)0 : mx(ab(ap), ab(am));
}

// ---- minmod-limited slope, linearreconstruction.jl:28-43 -------------------------------------------
// minmod(a,b,c) = min if all > 0, max if all < 0, else 0 -- evaluated without branches as
// max(0, min3) + min(0, max3): exactly one of the two terms is non-zero, so the value is identical.
template <class C>
__device__ __forceinline__ C minmod3(C a, C b, C c) {
    return mx((C)0, mn(mn(a, b), c)) + mn((C)0, mx(mx(a, b), c));
}
// ALT (fast Float64 only): the same value with 3 compare-and-selects instead of 6 --
//   sign(f) * min(theta * min(|f|,|b|), (|f|+|b|)/2) if f and b have the same sign, else 0
// (theta * x is monotone, so theta * min = min of the products; |f|+|b| = |f+b| for equal signs).  Round 1 selected it in
// stage 1 of the generic Float64 kernel; since round 2 the fast Float64 policy runs on kernels_fused_d1.cuh (whose slope()
// is this formula) and the generic kernels use ONE limiter form everywhere (ALT = false at every call site).
template <class P, bool ALT = false>
__device__ __forceinline__ typename P::C minmod_slope(typename P::S l, typename P::S c, typename P::S r,
                                                      typename P::C theta) {
    using S = typename P::S;
    using C = typename P::C;
    S f = r - c;
    S b = c - l;
    if constexpr (ALT && !P::strict && sizeof(C) == 8) {
        const C af = ab(f), abb = ab(b);
        const C m1 = af < abb ? af : abb;
        const C t = theta * m1, cen = (C)0.5 * (af + abb);
        const C m2 = t < cen ? t : cen;
        // same strict sign, tested on the sign bits: a product f * b would underflow to 0 for tiny differences
        const bool same = ((__double2hiint((double)f) ^ __double2hiint((double)b)) >= 0) && m1 > (C)0;
        return same ? cps(m2, (C)f) : (C)0;
    }
    C cen = (C)(f + b) * (C)0.5;   // (f+b)/2.0 : the S-precision sum, then an exact halving
    return minmod3<C>(theta * f, cen, theta * b);
}

// ---- reconstruction of one cell: face values (h, hu, hv) on its left/right face --------------------
// linearreconstruction.jl:73-100 (practical SWE: w-slope clamp, then subtract face bathymetry);
// :44-56 for Pure; noreconstruction.jl:30-46 for RECON == 0.
template <class P, int RECON, bool PURE, bool ALT = false>
__device__ __forceinline__ void recon_cell(const Par<typename P::C>& p, const typename P::S l[3],
                                           const typename P::S c[3], const typename P::S r[3],
                                           typename P::C Bl, typename P::C Br, typename P::C Bc,
                                           typename P::S L[3], typename P::S R[3]) {
    using S = typename P::S;
    using C = typename P::C;
    if constexpr (RECON == 0) {
        L[0] = R[0] = PURE ? c[0] : (S)(c[0] - Bc);
        L[1] = R[1] = c[1];
        L[2] = R[2] = c[2];
    } else {
        C s[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) s[k] = minmod_slope<P, ALT>(l[k], c[k], r[k], p.theta);
        const C half = (C)0.5;
        if constexpr (!PURE) {
            const S w = c[0];
            const bool c1 = w - half * s[0] < Bl;          // negative depth on the left face  (:81-84)
            const bool c2 = w + half * s[0] < Br;          // negative depth on the right face (:86-89)
            const C s1 = (C)2.0 * (w - Bl), s2 = (C)2.0 * (Br - w);
            s[0] = c1 ? s1 : (c2 ? s2 : s[0]);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            L[k] = (S)(c[k] - half * s[k]);
            R[k] = (S)(c[k] + half * s[k]);
        }
        if constexpr (!PURE) {
            L[0] = (S)(L[0] - Bl);
            R[0] = (S)(R[0] - Br);
        }
    }
}

// ---- bottom-slope source, source_term_bottom.jl:38 ---------------------------------------------------
//   out.mom += -g*((B_right - B_left)/dx)*((h_right + h_left)/2.0)
template <class P>
__device__ __forceinline__ typename P::C bottom_term(const Par<typename P::C>& p, typename P::C Bl,
                                                     typename P::C Br, typename P::S hl, typename P::S hr,
                                                     typename P::C delta, typename P::C rdelta) {
    using S = typename P::S;
    using C = typename P::C;
    S hsum = hr + hl;
    if constexpr (P::strict) return -p.g * ((Br - Bl) / delta) * (hsum / (C)2.0);
    else return -p.g * ((Br - Bl) * rdelta) * ((C)hsum * (C)0.5);
}

// ---- implicit friction, friction.jl:30-78: returns 1 - dt*c so that out.mom /= it -------------------
template <class P, bool PURE>
__device__ __forceinline__ typename P::C friction_denominator(const Par<typename P::C>& p, typename P::S pw,
                                                              typename P::S phu, typename P::S phv,
                                                              typename P::C Bm, typename P::S dt_s) {
    using C = typename P::C;
    C hstar;
    if constexpr (PURE) hstar = pw - (C)0.0;
    else hstar = desing_practical<P, C>(p, pw - Bm);
    C denom;
    if (p.friction == SWE_B200_FRICTION_BH2021) denom = hstar * hstar;
    else if (p.friction == SWE_B200_FRICTION_BSA2012) denom = cb(hstar) * hstar;
    else denom = cb(hstar) * hstar * hstar;
    if constexpr (P::strict) {
        C u = phu / hstar;
        C v = phv / hstar;
        C speed = sq(u * u + v * v);
        C fs = -p.Cz * speed / denom;
        return (C)1 - dt_s * fs;
    } else {      // |(hu, hv)| / (|h*| denom) with one reciprocal
        C speed = sq((C)phu * phu + (C)phv * phv);
        C fs = p.Cz * speed * rcp_fast(ab(hstar) * denom);
        return (C)1 + dt_s * fs;
    }
}

// ---- dry-cell clean-up, timestepper.jl:41-51 ---------------------------------------------------------
template <class P>
__device__ __forceinline__ void post_proc_cell(const Par<typename P::C>& p, typename P::S o[3],
                                               typename P::C Bm) {
    using S = typename P::S;
    using C = typename P::C;
    if (o[0] - Bm < p.cutoff) {
        o[0] = (S)mx((C)o[0], Bm);
        o[1] = (S)0;
        o[2] = (S)0;
    }
}

// ---- ghost-cell source map (bc.jl:29-40 periodic, :81-110 wall).  0-based index a in [0,N),
// returns the 0-based index whose value the BC copies into a (a itself if not a physical ghost)
// and whether the normal momentum changes sign. ------------------------------------------------------
__device__ __forceinline__ int bc_source(int a, int N, int gc, int bc, bool lo_phys, bool hi_phys, bool& flip, int dir = 0) {
    flip = false;
    if (a < gc) {
        if (!lo_phys) return a;
        if (bc == SWE_B200_BC_PERIODIC) return a + N - 2 * gc;
        if (bc == SWE_B200_BC_NEUMANN) return dir == 0 ? 2 * gc - 1 - a : a + N - 2 * gc;   // bc.jl:51,56 (X: mirror, Y: the
        flip = (bc == SWE_B200_BC_WALL);                                                     // image of the top interior rows)
        return 2 * gc - 1 - a;
    }
    if (a >= N - gc) {
        if (!hi_phys) return a;
        if (bc == SWE_B200_BC_PERIODIC) return a - N + 2 * gc;
        // NeumannBC high side (bc.jl:52-53, 57-58): ghost g=1 copies the outermost ghost, then g=2 copies g=1 back --
        // both end up holding the OLD outermost ghost cell, which never sees interior data and evolves on its own
        // (sources, friction, dry clean-up act on all cells): see k_neumann_dU / k_neumann_ghosts
        if (bc == SWE_B200_BC_NEUMANN) return N - 1;
        flip = (bc == SWE_B200_BC_WALL);
        return 2 * (N - gc) - 1 - a;
    }
    return a;
}

SWE_NS_END  // namespace swe

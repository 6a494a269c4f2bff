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
#include "kernels_fused.cuh"

namespace swe {
namespace x2 {

typedef float2 f2;

constexpr int VW = 60;       // valid (output) columns per warp: lanes 1..30, two columns each
constexpr int WPB = 4;       // warps per block
constexpr int ROWS = 64;     // rows marched per block

__device__ __forceinline__ f2 mk(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ f2 sp(float a) { return make_float2(a, a); }
__device__ __forceinline__ f2 add(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 mul(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 sub(f2 a, f2 b) { return __ffma2_rn(b, sp(-1.0f), a); }   // a - b
__device__ __forceinline__ f2 swp(f2 a) { return make_float2(a.y, a.x); }
__device__ __forceinline__ f2 neg(f2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float rcp_a(float x) {
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float sqrt_a(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float cbrt_a(float x) {   // sign(x) * 2^(log2|x| / 3): two MUFU
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
    float cutoff, kappa, constB;
    f2 theta, half, mhalf, quarter, g, K /*0.5 rho g*/, i2k, hk, rho, nrdx, nrdy, botx, boty, Cz;
};

struct Args2 {
    Par2 p;
    const float* U;
    const float* D;
    float* out;
    const float* B;
    const double* factor;
    Ctl* ctl;
};

// ---- minmod(a,b,c), linearreconstruction.jl:28-35, as the median of (min3, 0, max3) ----------------
__device__ __forceinline__ float minmod3f(float a, float b, float c) {
    const float lo = fminf(fminf(a, b), c), hi = fmaxf(fmaxf(a, b), c);
    return fminf(fmaxf(lo, 0.0f), hi);
}

// ---- reconstruction of a pair of cells from the centre values and the backward / forward differences
// (linearreconstruction.jl:37-43, 73-100; noreconstruction.jl:30-46).  tBl/tBr = 2 * face bathymetry.
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
    const f2 t = fma2(hh, p.i2k, p.hk);
    const f2 hs = mk(fmaxf(hm.x, fminf(t.x, p.kappa)), fmaxf(hm.y, fminf(t.y, p.kappa)));
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
// A warp whose faces are all dry on both sides skips the evaluation (all 32 lanes must be converged).
template <int DIR, bool SPEED>
__device__ __forceinline__ void cu2(const Par2& p, const f2 Um[3], const f2 Up[3], f2 F[3], f2& speed) {
    const bool any = Um[0].x > p.cutoff || Um[0].y > p.cutoff || Up[0].x > p.cutoff || Up[0].y > p.cutoff;
    if (!__any_sync(0xffffffffu, any)) {
        F[0] = F[1] = F[2] = sp(0.0f);
        speed = sp(0.0f);
        return;
    }
    f2 Fm[3], Fp[3], epm, enm, epp, enp;
    side2<DIR>(p, Um[0], Um[1], Um[2], Fm, epm, enm);
    side2<DIR>(p, Up[0], Up[1], Up[2], Fp, epp, enp);
    const f2 ap = mk(fmaxf(fmaxf(epm.x, epp.x), 0.0f), fmaxf(fmaxf(epm.y, epp.y), 0.0f));   //  a+
    const f2 an = mk(fmaxf(fmaxf(enm.x, enp.x), 0.0f), fmaxf(fmaxf(enm.y, enp.y), 0.0f));   // -a-
    const f2 d = add(ap, an);
    const bool okx = !(d.x < p.kappa), oky = !(d.y < p.kappa);                            // :67-69
    const f2 r = mk(okx ? rcp_a(d.x) : 0.0f, oky ? rcp_a(d.y) : 0.0f);
    const f2 wp = mul(ap, r), wn = mul(an, r);
    const f2 c2 = mul(ap, wn);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const f2 jn = sub(Um[k], Up[k]);
        f2 t = mul(wp, Fm[k]);
        t = fma2(wn, Fp[k], t);
        F[k] = fma2(c2, jn, t);
    }
    if (SPEED) speed = mk(okx ? fmaxf(ap.x, an.x) : 0.0f, oky ? fmaxf(ap.y, an.y) : 0.0f);
    else speed = sp(0.0f);
}

// ---- pointwise tail of a substep on a pair of cells: o = base + dt*d ; friction(prev) ; dry clean-up
// (rungekutta.jl:39-41, friction.jl:62-84, timestepper.jl:41-51) --------------------------------------------
__device__ __forceinline__ f2 friction_factor2(const Par2& p, const f2 prev[3], f2 Bc, f2 czdt) {
    const f2 h = sub(prev[0], Bc);
    const f2 t = fma2(mul(h, h), p.i2k, p.hk);
    const f2 hs = mk(copysignf(fmaxf(fabsf(h.x), fminf(t.x, p.kappa)), h.x),
                     copysignf(fmaxf(fabsf(h.y), fminf(t.y, p.kappa)), h.y));
    const f2 r = mk(rcp_a(hs.x), rcp_a(hs.y));
    const f2 u = mul(prev[1], r), v = mul(prev[2], r);
    const f2 s2 = fma2(u, u, mul(v, v));
    const f2 s = mk(sqrt_a(s2.x), sqrt_a(s2.y));
    f2 iden;   // 1 / (law denominator)
    if (p.friction == SWE_B200_FRICTION_BH2021) iden = mul(r, r);
    else {
        const f2 rc = mk(rcp_a(cbrt_a(hs.x)), rcp_a(cbrt_a(hs.y)));
        iden = mul(rc, r);
        if (p.friction != SWE_B200_FRICTION_BSA2012) iden = mul(iden, r);
    }
    const f2 den = fma2(mul(s, iden), czdt, sp(1.0f));     // 1 - dt*(-Cz*speed/denominator)
    return mk(rcp_a(den.x), rcp_a(den.y));
}
__device__ __forceinline__ void postproc2(const Par2& p, f2 o[3], f2 Bc) {
    const f2 dpt = sub(o[0], Bc);
    if (dpt.x < p.cutoff) { o[0].x = fmaxf(o[0].x, Bc.x); o[1].x = 0.0f; o[2].x = 0.0f; }
    if (dpt.y < p.cutoff) { o[0].y = fmaxf(o[0].y, Bc.y); o[1].y = 0.0f; o[2].y = 0.0f; }
}

struct CornerRow {
    f2 e;       // corners (i, r), (i+1, r)
    float e2;   // corner (i+2, r)
    f2 ty;      // 2 * B_face_y of the two cells above this corner row
};

template <int RECON, bool HASB, int STAGE>
__global__ void __launch_bounds__(WPB * 32, 4) k_stage_x2(const Args2 a) {
    const Par2& p = a.p;
    const int lane = threadIdx.x & 31;
    const int tile = blockIdx.x * WPB + (threadIdx.x >> 5);
    const int Nx = p.Nx, Ny = p.Ny;
    if (tile * VW + 2 >= Nx - 2) return;                 // warp-uniform: no output column
    const int i = tile * VW + 2 * lane;                  // first of this lane's two columns (even)
    const int ic = min(i, Nx - 2);                       // clamped for loads (Nx is even)
    const int ja = 2 + blockIdx.y * p.rows;
    const int jb = min(ja + p.rows, Ny - 2);
    const bool in_x = i <= Nx - 4;
    const bool out_pair = lane >= 1 && lane <= 30 && in_x;
    const bool facea_ok = out_pair;                      // face (i | i+1)   belongs to an interior cell
    const bool faceb_ok = lane <= 30 && in_x;            // face (i+1 | i+2)
    const long long plane = p.plane;
    const int Bp = p.Bp;

    f2 dt = sp(0.0f), czdt = sp(0.0f);
    int xs = ic;
    bool xswap = false, xflip = false;
    if (STAGE == 2) {
        const float dts = (float)a.ctl->dt_s;
        dt = sp(dts);
        czdt = mul(p.Cz, dt);
        if (ic < 2 || ic >= Nx - 2) {                    // a pair of x ghosts: where the BC copies them from
            const bool lo = ic < 2;
            if (p.bc == SWE_B200_BC_PERIODIC) xs = lo ? Nx - 4 : 2;
            else { xs = lo ? 2 : Nx - 4; xswap = true; xflip = true; }
        }
    }
    const float rain = (float)a.ctl->rain;
    const float infil_t = (float)a.ctl->infil;

    auto load_corner = [&](int row, CornerRow& c) {
        if (HASB) {
            const float* r = a.B + (size_t)row * Bp + ic;
            c.e = *reinterpret_cast<const f2*>(r);
            c.e2 = r[2];
        }
    };
    auto finish_corner = [&](CornerRow& c) {
        if (HASB) c.ty = mk(c.e.x + c.e.y, c.e.y + c.e2);
        else c.ty = sp(2.0f * p.constB);
    };
    auto bcell = [&](const CornerRow& lo, const CornerRow& hi) -> f2 {
        return HASB ? mul(p.quarter, add(lo.ty, hi.ty)) : sp(p.constB);
    };

    // raw global loads of row j (U^n, and dU in stage 2), at the BC source position in stage 2
    auto issue_row = [&](int j, f2 u[3], f2 d[3]) {
        if (STAGE == 1) {
            const float* s = a.U + (size_t)j * Nx + ic;
#pragma unroll
            for (int k = 0; k < 3; ++k) u[k] = *reinterpret_cast<const f2*>(s + k * plane);
        } else {
            bool yflip;
            const int ys = bc_source(j, Ny, 2, p.bc, p.ylo_phys, p.yhi_phys, yflip);
            const size_t off = (size_t)ys * Nx + xs;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                u[k] = *reinterpret_cast<const f2*>(a.U + off + k * plane);
                d[k] = *reinterpret_cast<const f2*>(a.D + off + k * plane);
            }
        }
    };
    // stage 1: q = U^n ; stage 2: q = U^1 = BC(postproc(friction(U^n + dt*dU))) rebuilt on the fly
    auto finish_row = [&](int j, const f2 u[3], const f2 d[3], f2 Bc_here, f2 q[3], f2 un[3]) {
        if (STAGE == 1) {
#pragma unroll
            for (int k = 0; k < 3; ++k) q[k] = u[k];
        } else {
            bool yflip;
            const int ys = bc_source(j, Ny, 2, p.bc, p.ylo_phys, p.yhi_phys, yflip);
#pragma unroll
            for (int k = 0; k < 3; ++k) un[k] = u[k];
            f2 Bm = Bc_here;
            if (HASB && (xs != ic || ys != j)) {         // ghost: bathymetry of the source cells
                const float* r0 = a.B + (size_t)ys * Bp + xs;
                const float m = r0[1] + r0[Bp + 1];
                Bm = mk(0.25f * (r0[0] + r0[Bp] + m), 0.25f * (m + r0[2] + r0[Bp + 2]));
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) q[k] = fma2(dt, d[k], u[k]);
            if (p.friction != SWE_B200_FRICTION_NONE) {
                const f2 fr = friction_factor2(p, u, Bm, czdt);
                q[1] = mul(q[1], fr);
                q[2] = mul(q[2], fr);
            }
            postproc2(p, q, Bm);
            if (xswap) {
#pragma unroll
                for (int k = 0; k < 3; ++k) q[k] = swp(q[k]);
            }
            if (xflip) q[1] = neg(q[1]);
            if (yflip) q[2] = neg(q[2]);
        }
    };

    // ---- prologue: state rows ja-2 .. ja+1, corner rows ja-2 .. ja+2 ----
    f2 q0[3], q1[3], q2[3], dprev[3];
    f2 un0[3], un1[3], un2[3];
    f2 Ry0[3], Gl[3], hsy;                                // upper face values / lower flux / (hL+hR) of cell j
    f2 msx = sp(0.0f), msy = sp(0.0f);
    CornerRow cr[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) { cr[r].e = sp(0.0f); cr[r].e2 = 0.0f; cr[r].ty = sp(0.0f); }
#pragma unroll
    for (int k = 0; k < 3; ++k) un0[k] = un1[k] = un2[k] = sp(0.0f);
    {
        CornerRow c0, c1, c2, c3, c4;                     // corner rows ja-2 .. ja+2
        load_corner(ja - 2, c0); load_corner(ja - 1, c1); load_corner(ja, c2); load_corner(ja + 1, c3);
        load_corner(ja + 2, c4);
        f2 u[4][3], d[4][3];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
            for (int k = 0; k < 3; ++k) d[r][k] = sp(0.0f);
            issue_row(ja - 2 + r, u[r], d[r]);
        }
        finish_corner(c0); finish_corner(c1); finish_corner(c2); finish_corner(c3); finish_corner(c4);
        f2 qa[3], qb[3], unx[3];
        finish_row(ja - 2, u[0], d[0], bcell(c0, c1), qa, unx);
        finish_row(ja - 1, u[1], d[1], bcell(c1, c2), qb, unx);
        finish_row(ja, u[2], d[2], bcell(c2, c3), q0, un0);
        finish_row(ja + 1, u[3], d[3], bcell(c3, c4), q1, un1);
        f2 b0[3], b1[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { b0[k] = sub(qb[k], qa[k]); b1[k] = sub(q0[k], qb[k]); dprev[k] = sub(q1[k], q0[k]); }
        f2 Lt[3], Rp[3], Ly0[3];
        recon2<RECON>(p, qb, b0, b1, c1.ty, c2.ty, bcell(c1, c2), Lt, Rp);          // cell ja-1: upper face only
        recon2<RECON>(p, q0, b1, dprev, c2.ty, c3.ty, bcell(c2, c3), Ly0, Ry0);     // cell ja
        f2 s;
        cu2<1, STAGE == 1>(p, Rp, Ly0, Gl, s);                                      // face ja-1/2
        if (STAGE == 1 && out_pair) msy = s;
        hsy = add(Ly0[0], Ry0[0]);
        cr[0] = c2; cr[1] = c3; cr[2] = c4;               // corner rows ja, ja+1, ja+2
    }

    // ---- march ----
    f2 ru[3], rd[3];
    CornerRow pc;
#pragma unroll
    for (int k = 0; k < 3; ++k) rd[k] = sp(0.0f);
    pc.e = sp(0.0f); pc.e2 = 0.0f; pc.ty = sp(0.0f);
    issue_row(min(ja + 2, Ny - 1), ru, rd);
    load_corner(min(ja + 3, Ny), pc);
#pragma unroll 2
    for (int j = ja; j < jb; ++j) {
        // window: q0 = row j, q1 = row j+1 ; corner rows cr[0..3] = j .. j+3
        cr[3] = pc;
        finish_corner(cr[3]);
        finish_row(j + 2, ru, rd, bcell(cr[2], cr[3]), q2, un2);
        issue_row(min(j + 3, Ny - 1), ru, rd);            // in flight during this row's arithmetic
        load_corner(min(j + 4, Ny), pc);

        f2 dU[3];
        // ================= x direction (row j) =================
        {
            f2 bwd[3], fwd[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const float ql = __shfl_up_sync(0xffffffffu, q0[k].y, 1);      // column i-1
                const float qr = __shfl_down_sync(0xffffffffu, q0[k].x, 1);    // column i+2
                const float dm = q0[k].y - q0[k].x;
                bwd[k] = mk(q0[k].x - ql, dm);
                fwd[k] = mk(dm, qr - q0[k].y);
            }
            f2 tBl, tBr;
            if (HASB) {
                tBl = add(cr[0].e, cr[1].e);
                tBr = mk(tBl.y, cr[0].e2 + cr[1].e2);
            } else tBl = tBr = sp(2.0f * p.constB);
            f2 Lx[3], Rx[3], Ln[3];
            recon2<RECON>(p, q0, bwd, fwd, tBl, tBr, bcell(cr[0], cr[1]), Lx, Rx);
#pragma unroll
            for (int k = 0; k < 3; ++k) Ln[k] = mk(Lx[k].y, __shfl_down_sync(0xffffffffu, Lx[k].x, 1));
            f2 F[3], s;
            cu2<0, STAGE == 1>(p, Rx, Ln, F, s);           // faces (i | i+1) and (i+1 | i+2)
            if (STAGE == 1) {
                if (facea_ok) msx.x = fmaxf(msx.x, s.x);
                if (faceb_ok) msx.y = fmaxf(msx.y, s.y);
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const f2 Fl = mk(__shfl_up_sync(0xffffffffu, F[k].y, 1), F[k].x);
                dU[k] = mul(p.nrdx, sub(F[k], Fl));
            }
            if (HASB && p.bottom_source) dU[1] = fma2(mul(p.botx, sub(tBr, tBl)), add(Lx[0], Rx[0]), dU[1]);
        }
        // ================= y direction (cell j, faces j-1/2 and j+1/2) =================
        {
            f2 fwd[3], Ly1[3], Ry1[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) fwd[k] = sub(q2[k], q1[k]);
            recon2<RECON>(p, q1, dprev, fwd, cr[1].ty, cr[2].ty, bcell(cr[1], cr[2]), Ly1, Ry1);   // cell j+1
            f2 G[3], s;
            cu2<1, STAGE == 1>(p, Ry0, Ly1, G, s);         // face j+1/2
            if (STAGE == 1 && out_pair) { msy.x = fmaxf(msy.x, s.x); msy.y = fmaxf(msy.y, s.y); }
#pragma unroll
            for (int k = 0; k < 3; ++k) dU[k] = fma2(p.nrdy, sub(G[k], Gl[k]), dU[k]);
            if (HASB && p.bottom_source) dU[2] = fma2(mul(p.boty, sub(cr[1].ty, cr[0].ty)), hsy, dU[2]);
#pragma unroll
            for (int k = 0; k < 3; ++k) { dprev[k] = fwd[k]; Ry0[k] = Ry1[k]; Gl[k] = G[k]; }
            hsy = add(Ly1[0], Ry1[0]);
        }
        // ================= per-cell sources (system.jl:71-73, list order) =================
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            const bool do_infil = (pass == 0) == (p.infil_first != 0);
            if (do_infil && p.infil_kind != SWE_B200_INFIL_NONE) {
                f2 f;
                if (a.factor) {
                    const double2 fc = *reinterpret_cast<const double2*>(a.factor + (size_t)j * Nx + ic);
                    f = mul(mk((float)fc.x, (float)fc.y), sp(infil_t));
                } else f = sp(infil_t);
                dU[0] = sub(dU[0], f);
            }
            if (!do_infil && p.rain_kind != SWE_B200_RAIN_NONE) dU[0] = add(dU[0], sp(rain));
        }
        // ================= store =================
        if (STAGE == 1) {
            if (out_pair) {
                float* o = a.out + (size_t)j * Nx + i;
#pragma unroll
                for (int k = 0; k < 3; ++k) *reinterpret_cast<f2*>(o + k * plane) = dU[k];
            }
        } else {
            // rungekutta.jl:39-49 stage 2, then friction (prev = U^1) and the dry clean-up
            f2 o[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) o[k] = mul(p.half, add(un0[k], fma2(dt, dU[k], q0[k])));
            const f2 Bm = bcell(cr[0], cr[1]);
            if (p.friction != SWE_B200_FRICTION_NONE) {
                const f2 fr = friction_factor2(p, q0, Bm, czdt);
                o[1] = mul(o[1], fr);
                o[2] = mul(o[2], fr);
            }
            postproc2(p, o, Bm);
            if (out_pair) {
                float* w = a.out + (size_t)j * Nx + i;
#pragma unroll
                for (int k = 0; k < 3; ++k) *reinterpret_cast<f2*>(w + k * plane) = o[k];
            }
        }
        // ================= rotate =================
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            q0[k] = q1[k]; q1[k] = q2[k];
            un0[k] = un1[k]; un1[k] = un2[k];
        }
        cr[0] = cr[1]; cr[1] = cr[2]; cr[2] = cr[3];
    }

    if (STAGE == 1) {
        // CFL reduction: warp shuffle -> one atomicMax per warp and direction (non-negative bit patterns)
        unsigned bx = __float_as_uint(fmaxf(msx.x, msx.y)), by = __float_as_uint(fmaxf(msy.x, msy.y));
        bx = __reduce_max_sync(0xffffffffu, bx);
        by = __reduce_max_sync(0xffffffffu, by);
        if (lane == 0) {
            if (bx) atomicMax(&a.ctl->maxspeed_bits[0], (unsigned long long)bx);
            if (by) atomicMax(&a.ctl->maxspeed_bits[1], (unsigned long long)by);
        }
    }
}

}  // namespace x2
}  // namespace swe

// f4_prototype.cu -- SURVEY 8(f4) feasibility: ONE sweep doing stage 2 of step n and stage 1 of step n+1
// (K2(n) + K1(n+1): U^{n+1} is consumed from registers by the next step's flux operator instead of being re-read).
//
// COMPILE-ONLY measurement (scratch/f4_measure.sh): registers, spills and static instructions per row of the fused sweep
// against K1 and K2 of the generic one-column kernel (kernels_fused.cuh), same policy, same flags.  It is NOT a
// product kernel and was never run: the prologue of the second pipeline, the boundary conditions of U^{n+1} (its ghost
// cells would have to be filled inside the sweep before the second operator reads them), the strip exchange and the
// CFL hand-over are left out -- the steady-state row body is complete, which is what the numbers are about.
//
// The sweep, for output row j of stage 2 (U^{n+1}(j) = o), runs the flux operator of step n+1 on row jj = j-2:
// its x stencil needs U^{n+1}(i-2..i+2, jj) (neighbour lanes: the 28 valid columns of a warp shrink to 24), its y
// stencil rows jj-1..jj+2, i.e. a second 3-row window (w0 = row j-2, w1 = row j-1, o = row j), a second set of carried
// face values / lower flux, and two more bathymetry corner rows.
#define SWE_FAST_MINMAX 1
#define SWE_FAST_DIV64 1
#define SWE_TU_NS tu_f4
#include "../include/swe_b200.h"
#include "../volumefluxes.jl_b200/csrc/kernels_fused.cuh"

SWE_NS_BEGIN

template <class P>
struct Stage23Args {
    StageArgs<P> s;                 // stage 2 of step n: U = U^n, D = dU^n, out = U^{n+1}
    typename P::S* out2;            // dU^{n+1} = L(U^{n+1})
    const typename P::S* carry;     // (prototype) where the second pipeline's carried state would come from at a chunk start
};

template <class P, int RECON, bool HASB>
__global__ void __launch_bounds__(FUSED_WPB * 32) k_stage23(const Stage23Args<P> aa) {
    using S = typename P::S;
    using C = typename P::C;
    const StageArgs<P>& a = aa.s;
    const Par<C>& p = a.p;
    constexpr int STAGE = 2;
    const int lane = threadIdx.x & 31;
    const int wcol = blockIdx.x * FUSED_WPB + (threadIdx.x >> 5);
    const int Nx = p.Nx, Ny = p.Ny;
    constexpr int VW2 = 24;                                // valid columns of the SECOND operator per warp
    if (wcol * VW2 + 4 >= Nx - 2) return;
    const int i = wcol * VW2 + lane;
    const int ic = min(i, Nx - 1);
    const int ja = 2 + blockIdx.y * a.rows;
    const int jb = min(ja + a.rows, Ny - 2);
    const bool out_lane = lane >= 2 && lane <= 29 && i < Nx - 2;        // U^{n+1}
    const bool out2_lane = lane >= 4 && lane <= 27 && i < Nx - 4;       // dU^{n+1}
    const bool xface2_lane = lane >= 3 && lane <= 27 && i >= 1 && i <= Nx - 3;
    const long long plane = p.plane;
    const int Bp = a.Bp;

    C dt = (C)a.ctl->dt;
    S dt_s = (S)a.ctl->dt_s;
    bool xflip = false;
    const int xs = bc_source(ic, Nx, 2, p.bc, true, true, xflip);
    const C rain = (C)a.ctl->rain;
    const double infil_t = a.ctl->infil;

    S ca[4] = {0, 0, 0, 0}, cb[4] = {0, 0, 0, 0};
    S cam2 = 0, cbm2 = 0, cam1 = 0, cbm1 = 0;              // corner rows j-2, j-1 (second operator)
    auto load_corners = [&](int row, S& x, S& y) {
        if (HASB) {
            const S* r = a.B + (size_t)row * Bp + ic;
            x = r[0];
            y = r[1];
        }
    };
    auto bcell_of = [&](int r) -> C { return HASB ? (C)0.25 * (ca[r] + cb[r] + ca[r + 1] + cb[r + 1]) : p.constB; };
    auto bface_y = [&](int r) -> C { return HASB ? (C)0.5 * (ca[r] + cb[r]) : p.constB; };

    auto issue_row = [&](int j, S u[3], S d[3]) {
        bool yflip;
        const int ys = bc_source(j, Ny, 2, p.bc, p.ylo_phys, p.yhi_phys, yflip, 1);
        const size_t off = (size_t)ys * Nx + xs;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            u[k] = a.U[off + k * plane];
            d[k] = a.D[off + k * plane];
        }
    };
    auto finish_row = [&](int j, const S u[3], const S d[3], C bc_actual, S q[3], S un[3]) {
        bool yflip;
        const int ys = bc_source(j, Ny, 2, p.bc, p.ylo_phys, p.yhi_phys, yflip, 1);
#pragma unroll
        for (int k = 0; k < 3; ++k) un[k] = u[k];
        C Bm = bc_actual;
        const S* wr = !HASB ? nullptr : (j < 2 ? a.wrapLo : (j >= Ny - 2 ? a.wrapHi : nullptr));
        if (HASB && (xs != ic || ys != j || wr)) {
            const S* r0 = wr ? wr + (size_t)(j < 2 ? j : j - (Ny - 2)) * Bp + xs : a.B + (size_t)ys * Bp + xs;
            Bm = (C)0.25 * (r0[0] + r0[1] + r0[Bp] + r0[Bp + 1]);
        }
        substep_tail<P>(p, u, d, dt, dt_s, Bm, q);
        if (xflip) q[1] = -q[1];
        if (yflip) q[2] = -q[2];
    };
    auto load_row = [&](int j, C bc_actual, S q[3], S un[3]) {
        S u[3], d[3] = {0, 0, 0};
        issue_row(j, u, d);
        finish_row(j, u, d, bc_actual, q, un);
    };

    // ---- prologue of the first pipeline: as k_stage<.., STAGE = 2> ----
    S qm[3], q0[3], q1[3], q2[3];
    S un0[3] = {0, 0, 0}, un1[3] = {0, 0, 0}, un2[3] = {0, 0, 0}, unx[3];
    S Ly0[3], Ry0[3];
    C Gl[3];
    {
        load_corners(ja - 2, ca[0], cb[0]);
        load_corners(ja - 1, ca[1], cb[1]);
        load_corners(ja, ca[2], cb[2]);
        load_corners(ja + 1, ca[3], cb[3]);
        cam2 = ca[0]; cbm2 = cb[0]; cam1 = ca[1]; cbm1 = cb[1];
        load_row(ja - 2, bcell_of(0), qm, unx);
        load_row(ja - 1, bcell_of(1), q0, unx);
        load_row(ja, bcell_of(2), q1, un0);
        S Lt[3], Rp[3];
        recon_cell<P, RECON, false, false>(p, qm, q0, q1, bface_y(1), bface_y(2), bcell_of(1), Lt, Rp);
        ca[0] = ca[1]; cb[0] = cb[1]; ca[1] = ca[2]; cb[1] = cb[2]; ca[2] = ca[3]; cb[2] = cb[3];
        load_corners(ja + 2, ca[3], cb[3]);
        load_row(ja + 1, bcell_of(2), q2, un1);
        recon_cell<P, RECON, false, false>(p, q0, q1, q2, bface_y(1), bface_y(2), bcell_of(1), Ly0, Ry0);
        C s;
        central_upwind<P, false, 1, true>(p, Rp, Ly0, Gl, s);
#pragma unroll
        for (int k = 0; k < 3; ++k) { qm[k] = q0[k]; q0[k] = q1[k]; q1[k] = q2[k]; }
        ca[0] = ca[1]; cb[0] = cb[1]; ca[1] = ca[2]; cb[1] = cb[2]; ca[2] = ca[3]; cb[2] = cb[3];
    }
    // ---- (prototype) state of the second pipeline at the chunk start: rows ja-2, ja-1 of U^{n+1}, the face values of
    // cell ja-2 and the flux through face ja-5/2 -- a real kernel recomputes them (two more rows of stage 2) ----
    S w0[3], w1[3], Ly0n[3], Ry0n[3];
    C Gln[3];
    S msx = 0, msy = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const S* c = aa.carry + (size_t)ja * Nx + ic + k * plane;
        w0[k] = c[0]; w1[k] = c[Nx]; Ly0n[k] = c[2 * Nx]; Ry0n[k] = c[3 * Nx]; Gln[k] = (C)c[4 * Nx];
    }

    S ru[3], rd[3] = {0, 0, 0}, pa = 0, pb = 0;
    issue_row(min(ja + 2, Ny - 1), ru, rd);
    load_corners(min(ja + 3, Ny), pa, pb);
#pragma unroll FUSED_UNROLL_N
    for (int j = ja; j < jb; ++j) {
        ca[3] = pa;
        cb[3] = pb;
        finish_row(j + 2, ru, rd, bcell_of(2), q2, un2);
        issue_row(min(j + 3, Ny - 1), ru, rd);
        load_corners(min(j + 4, Ny), pa, pb);

        // ======================= stage 2 of step n, row j (identical to k_stage) =======================
        S dU[3];
        {
            S ql[3], qr[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) { ql[k] = shfl_up1(q0[k]); qr[k] = shfl_dn1(q0[k]); }
            const C Blx = HASB ? (C)0.5 * (ca[0] + ca[1]) : p.constB;
            const C Brx = HASB ? (C)0.5 * (cb[0] + cb[1]) : p.constB;
            S Lx[3], Rx[3], Ln[3];
            recon_cell<P, RECON, false, false>(p, ql, q0, qr, Blx, Brx, bcell_of(0), Lx, Rx);
#pragma unroll
            for (int k = 0; k < 3; ++k) Ln[k] = shfl_dn1(Lx[k]);
            C F[3], s;
            central_upwind<P, false, 0, true>(p, Rx, Ln, F, s);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const C Fl = shfl_up1(F[k]);
                dU[k] = (S)(-(p.rdx * (F[k] - Fl)));
            }
            if (HASB && p.bottom_source) dU[1] = (S)(dU[1] + bottom_term<P>(p, Blx, Brx, Lx[0], Rx[0], p.dx, p.rdx));
        }
        {
            S Ly1[3], Ry1[3];
            recon_cell<P, RECON, false, false>(p, q0, q1, q2, bface_y(1), bface_y(2), bcell_of(1), Ly1, Ry1);
            C G[3], s;
            central_upwind<P, false, 1, true>(p, Ry0, Ly1, G, s);
#pragma unroll
            for (int k = 0; k < 3; ++k) dU[k] = (S)(dU[k] - p.rdy * (G[k] - Gl[k]));
            if (HASB && p.bottom_source)
                dU[2] = (S)(dU[2] + bottom_term<P>(p, bface_y(0), bface_y(1), Ly0[0], Ry0[0], p.dy, p.rdy));
#pragma unroll
            for (int k = 0; k < 3; ++k) { Ly0[k] = Ly1[k]; Ry0[k] = Ry1[k]; Gl[k] = G[k]; }
        }
        auto sources = [&](int row, S d3[3]) {
#pragma unroll
            for (int pass = 0; pass < 2; ++pass) {
                const bool do_infil = (pass == 0) == (a.infil_first != 0);
                if (do_infil && a.infil_kind != SWE_B200_INFIL_NONE) {
                    C f;
                    if (a.factor) f = (C)(a.factor[(size_t)row * Nx + ic] * infil_t);
                    else f = (C)infil_t;
                    d3[0] = (S)(d3[0] - f);
                }
                if (!do_infil && a.rain_kind != SWE_B200_RAIN_NONE)
                    d3[0] = (S)(d3[0] + (a.rainp ? (C)a.rainp[(size_t)row * Nx + ic] : rain));
            }
        };
        sources(j, dU);
        S o[3];
        {
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                o[k] = (S)(q0[k] + dt_s * dU[k]);
                o[k] = (S)((C)0.5 * (S)(un0[k] + o[k]));
            }
            const C Bm = bcell_of(0);
            if (p.friction != SWE_B200_FRICTION_NONE) {
                const C den = friction_denominator<P, false>(p, q0[0], q0[1], q0[2], Bm, dt_s);
                const C rden = rcp_fast(den);
                o[1] = (S)(o[1] * rden);
                o[2] = (S)(o[2] * rden);
            }
            post_proc_cell<P>(p, o, Bm);
            if (out_lane) {
                S* w = a.out + (size_t)j * Nx + i;
                w[0] = o[0];
                w[plane] = o[1];
                w[2 * plane] = o[2];
            }
        }
        // ======================= stage 1 of step n+1, row jj = j-2 (window w0, w1, o) =======================
        {
            S dN[3];
            S ql[3], qr[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) { ql[k] = shfl_up1(w0[k]); qr[k] = shfl_dn1(w0[k]); }
            const C Blx = HASB ? (C)0.5 * (cam2 + cam1) : p.constB;
            const C Brx = HASB ? (C)0.5 * (cbm2 + cbm1) : p.constB;
            const C Bc0 = HASB ? (C)0.25 * (cam2 + cbm2 + cam1 + cbm1) : p.constB;          // cell j-2
            const C Bc1 = HASB ? (C)0.25 * (cam1 + cbm1 + ca[0] + cb[0]) : p.constB;        // cell j-1
            const C Bfm2 = HASB ? (C)0.5 * (cam2 + cbm2) : p.constB, Bfm1 = HASB ? (C)0.5 * (cam1 + cbm1) : p.constB;
            S Lx[3], Rx[3], Ln[3];
            recon_cell<P, RECON, false, false>(p, ql, w0, qr, Blx, Brx, Bc0, Lx, Rx);
#pragma unroll
            for (int k = 0; k < 3; ++k) Ln[k] = shfl_dn1(Lx[k]);
            C F[3], s;
            central_upwind<P, false, 0, true>(p, Rx, Ln, F, s);
            if (xface2_lane) msx = (S)mx((C)msx, s);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const C Fl = shfl_up1(F[k]);
                dN[k] = (S)(-(p.rdx * (F[k] - Fl)));
            }
            if (HASB && p.bottom_source) dN[1] = (S)(dN[1] + bottom_term<P>(p, Blx, Brx, Lx[0], Rx[0], p.dx, p.rdx));
            S Ly1[3], Ry1[3];
            recon_cell<P, RECON, false, false>(p, w0, w1, o, Bfm1, bface_y(0), Bc1, Ly1, Ry1);     // cell j-1
            C G[3];
            central_upwind<P, false, 1, true>(p, Ry0n, Ly1, G, s);                               // face j-3/2
            if (out2_lane) msy = (S)mx((C)msy, s);
#pragma unroll
            for (int k = 0; k < 3; ++k) dN[k] = (S)(dN[k] - p.rdy * (G[k] - Gln[k]));
            if (HASB && p.bottom_source) dN[2] = (S)(dN[2] + bottom_term<P>(p, Bfm2, Bfm1, Ly0n[0], Ry0n[0], p.dy, p.rdy));
#pragma unroll
            for (int k = 0; k < 3; ++k) { Ly0n[k] = Ly1[k]; Ry0n[k] = Ry1[k]; Gln[k] = G[k]; }
            sources(j - 2, dN);
            if (out2_lane) {
                S* w = aa.out2 + (size_t)(j - 2) * Nx + i;
                w[0] = dN[0];
                w[plane] = dN[1];
                w[2 * plane] = dN[2];
            }
        }
        // ======================= rotate both pipelines =======================
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            qm[k] = q0[k]; q0[k] = q1[k]; q1[k] = q2[k];
            un0[k] = un1[k]; un1[k] = un2[k];
            w0[k] = w1[k]; w1[k] = o[k];
        }
        cam2 = cam1; cbm2 = cbm1; cam1 = ca[0]; cbm1 = cb[0];
        ca[0] = ca[1]; cb[0] = cb[1]; ca[1] = ca[2]; cb[1] = cb[2]; ca[2] = ca[3]; cb[2] = cb[3];
    }
    {
        unsigned long long bx = Bits<S>::enc(msx), by = Bits<S>::enc(msy);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long tx = __shfl_xor_sync(0xffffffffu, bx, o);
            unsigned long long ty = __shfl_xor_sync(0xffffffffu, by, o);
            bx = tx > bx ? tx : bx;
            by = ty > by ? ty : by;
        }
        if (lane == 0) {
            if (bx) atomicMax(&a.ctl->maxspeed_bits[0], bx);
            if (by) atomicMax(&a.ctl->maxspeed_bits[1], by);
        }
    }
}

// the three kernels to compare, same policy (native Float32, as engine_fast.cu compiles it)
using PF = Policy<float, float, false>;
template __global__ void k_stage<PF, 1, true, 1>(const StageArgs<PF>);
template __global__ void k_stage<PF, 1, true, 2>(const StageArgs<PF>);
template __global__ void k_stage23<PF, 1, true>(const Stage23Args<PF>);
using PD = Policy<double, double, false>;
template __global__ void k_stage<PD, 1, true, 1>(const StageArgs<PD>);
template __global__ void k_stage<PD, 1, true, 2>(const StageArgs<PD>);
template __global__ void k_stage23<PD, 1, true>(const Stage23Args<PD>);

SWE_NS_END

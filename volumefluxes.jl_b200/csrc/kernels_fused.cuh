// kernels_fused.cuh -- the fused sm_100a stage kernels (the hot path).
//
// One RK2 step = K1 (stage 1) -> k_ctl_dt -> K2 (stage 2) -> k_tail (kernels_exchange.cuh):
//   K1  reads U^n (3 words) + B (1)            writes dU = L(U^n) (3)      + CFL max (atomics)
//   K2  reads U^n (3) + dU (3) + B (1)         writes U^{n+1} (3)
// = 17 words per cell per step instead of the reference's ~183.  K2 rebuilds
// U^1 = BC(postproc(friction(U^n + dt*dU))) on the fly for its whole stencil window, evaluates
// L(U^1), and finishes the Heun average, friction and dry clean-up in registers.
//
// Mapping: each thread owns ONE column i and marches over a chunk of ROWS rows.
//   * y-direction: a 4-row sliding window of the state lives in registers; the face values of the
//     current cell and the flux through its lower face are carried from the previous row, so every
//     y-face is reconstructed and evaluated exactly once;
//   * x-direction: neighbours are other lanes of the same warp -> warp shuffles (no shared memory,
//     no block barrier).  A warp covers 32 consecutive columns and produces the inner 28.
// Loads are coalesced 128 B rows per warp; nothing is re-read from DRAM (the 4-column overlap of
// adjacent warps and the 2+2 halo rows of a chunk hit in L1/L2).
//
// Reference sequence being fused (paths under the reference's src/): timestepper/rungekutta.jl:26-52,
// system.jl:62-75, reconstruction/linearreconstruction.jl:58-102, numericalflux/numericalflux.jl:21-33,
// sourceterms/*.jl, friction.jl:80-84, timestepper/timestepper.jl:41-51, bc.jl:29-40,81-110.
#pragma once
#include "kernels_unfused.cuh"

SWE_NS_BEGIN

constexpr int FUSED_VW = 28;     // valid (output) columns per warp
constexpr int FUSED_WPB = 4;     // warps per block
constexpr int FUSED_ROWS = 64;   // rows marched per block
// unroll of the row march: the register windows (state rows, U^n rows, corner rows) rotate with period 3
#ifndef FUSED_UNROLL
#define FUSED_UNROLL 2
#endif
constexpr int FUSED_UNROLL_N = FUSED_UNROLL;

template <class P>
struct StageArgs {
    using S = typename P::S;
    Par<typename P::C> p;
    const S* U;        // U^n, ghosts valid
    const S* D;        // dU (stage 2 input)
    S* out;            // stage 1: dU ; stage 2: U^{n+1}
    const S* B;        // corners (pitch Bp) or nullptr
    const S* wrapLo;   // y-strips whose low / high ghost rows are a wrap-around image (periodic, NeumannBC low side): the gc+1
    const S* wrapHi;   // corner rows (pitch Bp) of the image's SOURCE cell rows; nullptr otherwise
    int Bp;
    const double* factor;   // Horton factor plane or nullptr
    const double* rainp;    // FunctionalRain values per cell or nullptr
    Ctl* ctl;
    int rain_kind, infil_kind, infil_first;
    int rows;          // rows per chunk
};

template <class T>
__device__ __forceinline__ T shfl_up1(T v) { return __shfl_up_sync(0xffffffffu, v, 1); }
template <class T>
__device__ __forceinline__ T shfl_dn1(T v) { return __shfl_down_sync(0xffffffffu, v, 1); }

// pointwise tail of a substep on one cell: out = state + dt*d ; friction(prev = state) ; post-proc
// (rungekutta.jl:39-41 / forwardeuler.jl:42-44, friction.jl:80-84, timestepper.jl:41-51)
template <class P>
__device__ __forceinline__ void substep_tail(const Par<typename P::C>& p, const typename P::S st[3],
                                             const typename P::S d[3], typename P::C dt, typename P::S dt_s,
                                             typename P::C Bm, typename P::S o[3]) {
    using S = typename P::S;
    using C = typename P::C;
#pragma unroll
    for (int k = 0; k < 3; ++k) o[k] = (S)(st[k] + dt * d[k]);
    if (p.friction != SWE_B200_FRICTION_NONE) {
        const C den = friction_denominator<P, false>(p, st[0], st[1], st[2], Bm, dt_s);
        if constexpr (P::strict) {
            o[1] = (S)(o[1] / den);
            o[2] = (S)(o[2] / den);
        } else {
            const C rden = rcp_fast(den);
            o[1] = (S)(o[1] * rden);
            o[2] = (S)(o[2] * rden);
        }
    }
    post_proc_cell<P>(p, o, Bm);
}

template <class P, int RECON, bool HASB, int STAGE>
__global__ void __launch_bounds__(FUSED_WPB * 32) k_stage(const StageArgs<P> a) {
    using S = typename P::S;
    using C = typename P::C;
    const Par<C>& p = a.p;
    const int lane = threadIdx.x & 31;
    const int wcol = blockIdx.x * FUSED_WPB + (threadIdx.x >> 5);
    const int Nx = p.Nx, Ny = p.Ny;
    if (wcol * FUSED_VW + 2 >= Nx - 2) return;            // warp-uniform: no output column
    const int i = wcol * FUSED_VW + lane;                 // this lane's column (0-based, total index)
    const int ic = min(i, Nx - 1);                        // clamped for loads
    const int ja = 2 + blockIdx.y * a.rows;
    const int jb = min(ja + a.rows, Ny - 2);
    const bool out_lane = lane >= 2 && lane <= 29 && i < Nx - 2;
    const bool xface_lane = lane >= 1 && lane <= 29 && i >= 1 && i <= Nx - 3;
    const long long plane = p.plane;
    const int Bp = a.Bp;

    // stage-2 control scalars and x ghost map
    C dt = 0;
    S dt_s = 0;
    int xs = ic;
    bool xflip = false;
    if (STAGE == 2) {
        dt = (C)a.ctl->dt;
        dt_s = (S)a.ctl->dt_s;
        xs = bc_source(ic, Nx, 2, p.bc, true, true, xflip);
    }
    const C rain = (C)a.ctl->rain;
    const double infil_t = a.ctl->infil;

    // ---- bathymetry corner rows: ca[r], cb[r] = B[ic, j+r], B[ic+1, j+r] ----
    S ca[4] = {0, 0, 0, 0}, cb[4] = {0, 0, 0, 0};
    auto load_corners = [&](int row, S& x, S& y) {
        if (HASB) {
            const S* r = a.B + (size_t)row * Bp + ic;
            x = r[0];
            y = r[1];
        }
    };
    auto bcell_of = [&](int r) -> C {   // B_cell of the cell whose lower corner row is ca[r]
        return HASB ? (C)0.25 * (ca[r] + cb[r] + ca[r + 1] + cb[r + 1]) : p.constB;
    };
    auto bface_y = [&](int r) -> C { return HASB ? (C)0.5 * (ca[r] + cb[r]) : p.constB; };

    // ---- row loader, split in two so that the loads of row j+3 are in flight during row j's math:
    //      issue_row : raw global loads of U^n (and dU in stage 2) at (column ic, row j)
    //      finish_row: stage 1: q = U^n ; stage 2: q = the rebuilt U^1 = BC(postproc(friction(U^n + dt*dU)))
    auto issue_row = [&](int j, S u[3], S d[3]) {
        if (STAGE == 1) {
            const S* s = a.U + (size_t)j * Nx + ic;
            u[0] = s[0];
            u[1] = s[plane];
            u[2] = s[2 * plane];
        } else {
            bool yflip;
            const int ys = bc_source(j, Ny, 2, p.bc, p.ylo_phys, p.yhi_phys, yflip, 1);
            const size_t off = (size_t)ys * Nx + xs;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                u[k] = a.U[off + k * plane];
                d[k] = a.D[off + k * plane];
            }
        }
    };
    auto finish_row = [&](int j, const S u[3], const S d[3], C bc_actual, S q[3], S un[3]) {
        if (STAGE == 1) {
#pragma unroll
            for (int k = 0; k < 3; ++k) q[k] = u[k];
        } else {
            bool yflip;
            const int ys = bc_source(j, Ny, 2, p.bc, p.ylo_phys, p.yhi_phys, yflip, 1);
#pragma unroll
            for (int k = 0; k < 3; ++k) un[k] = u[k];
            C Bm = bc_actual;
            // a ghost cell is rebuilt with the bathymetry of the cell the BC copies it from: (xs, ys) of this strip, or --
            // where the ghost rows came from the far end of the domain by exchange -- the wrap rows handed over by the host
            const S* wr = !HASB ? nullptr : (j < 2 ? a.wrapLo : (j >= Ny - 2 ? a.wrapHi : nullptr));
            if (HASB && (xs != ic || ys != j || wr)) {
                const S* r0 = wr ? wr + (size_t)(j < 2 ? j : j - (Ny - 2)) * Bp + xs : a.B + (size_t)ys * Bp + xs;
                Bm = (C)0.25 * (r0[0] + r0[1] + r0[Bp] + r0[Bp + 1]);
            }
            substep_tail<P>(p, u, d, dt, dt_s, Bm, q);
            if (xflip) q[1] = -q[1];
            if (yflip) q[2] = -q[2];
        }
    };
    auto load_row = [&](int j, C bc_actual, S q[3], S un[3]) {
        S u[3], d[3] = {0, 0, 0};
        issue_row(j, u, d);
        finish_row(j, u, d, bc_actual, q, un);
    };

    // ---- prologue: rows ja-2 .. ja+1, corner rows ja-1 .. ja+1(+1) ----
    S qm[3], q0[3], q1[3], q2[3];
    S un0[3] = {0, 0, 0}, un1[3] = {0, 0, 0}, un2[3] = {0, 0, 0}, unx[3];
    S Ly0[3], Ry0[3];
    C Gl[3];
    S msx = 0, msy = 0;
    {
        // corner rows ja-2 .. ja+1 into ca[0..3]
        load_corners(ja - 2, ca[0], cb[0]);
        load_corners(ja - 1, ca[1], cb[1]);
        load_corners(ja, ca[2], cb[2]);
        load_corners(ja + 1, ca[3], cb[3]);
        load_row(ja - 2, bcell_of(0), qm, unx);
        load_row(ja - 1, bcell_of(1), q0, unx);
        load_row(ja, bcell_of(2), q1, un0);
        S Lt[3], Rp[3];
        // cell ja-1 (only its right/upper face value is needed)
        recon_cell<P, RECON, false, false>(p, qm, q0, q1, bface_y(1), bface_y(2), bcell_of(1), Lt, Rp);
        // shift corners so that ca[0] = row ja-1, then load row ja+2 -> needed for bcell(ja+1)
        ca[0] = ca[1]; cb[0] = cb[1]; ca[1] = ca[2]; cb[1] = cb[2]; ca[2] = ca[3]; cb[2] = cb[3];
        load_corners(ja + 2, ca[3], cb[3]);
        load_row(ja + 1, bcell_of(2), q2, un1);
        // cell ja
        recon_cell<P, RECON, false, false>(p, q0, q1, q2, bface_y(1), bface_y(2), bcell_of(1), Ly0, Ry0);
        C s;
        central_upwind<P, false, 1, true>(p, Rp, Ly0, Gl, s);
        if (out_lane) msy = (S)s;
        // rotate the state window: q0 = row ja, q1 = row ja+1
#pragma unroll
        for (int k = 0; k < 3; ++k) { qm[k] = q0[k]; q0[k] = q1[k]; q1[k] = q2[k]; }
        // corners: ca[0] = row ja, ca[1] = row ja+1, ca[2] = row ja+2
        ca[0] = ca[1]; cb[0] = cb[1]; ca[1] = ca[2]; cb[1] = cb[2]; ca[2] = ca[3]; cb[2] = cb[3];
    }

    // ---- march ----
    S ru[3], rd[3] = {0, 0, 0}, pa = 0, pb = 0;          // prefetched raw row j+2 and corner row j+3
    issue_row(min(ja + 2, Ny - 1), ru, rd);
    load_corners(min(ja + 3, Ny), pa, pb);
#pragma unroll FUSED_UNROLL_N
    for (int j = ja; j < jb; ++j) {
        // window: qm = row j-1, q0 = row j, q1 = row j+1 ; corners ca[0..2] = rows j..j+2
        ca[3] = pa;                                        // corner row j+3: bcell(j+2) (stage 2 / RECON 0)
        cb[3] = pb;
        finish_row(j + 2, ru, rd, bcell_of(2), q2, un2);
        issue_row(min(j + 3, Ny - 1), ru, rd);             // in flight during this row's arithmetic
        load_corners(min(j + 4, Ny), pa, pb);

        S dU[3];
        // ================= x direction (row j) =================
        {
            S ql[3], qr[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) { ql[k] = shfl_up1(q0[k]); qr[k] = shfl_dn1(q0[k]); }
            const C Blx = HASB ? (C)0.5 * (ca[0] + ca[1]) : p.constB;
            const C Brx = HASB ? (C)0.5 * (cb[0] + cb[1]) : p.constB;
            S Lx[3], Rx[3], Ln[3];
            recon_cell<P, RECON, false, false>(p, ql, q0, qr, Blx, Brx, bcell_of(0), Lx, Rx);
#pragma unroll
            for (int k = 0; k < 3; ++k) Ln[k] = shfl_dn1(Lx[k]);      // left face value of cell i+1
            C F[3], s;
            central_upwind<P, false, 0, true>(p, Rx, Ln, F, s);             // face i+1/2
            if (STAGE == 1 && xface_lane) msx = (S)mx((C)msx, s);
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const C Fl = shfl_up1(F[k]);                           // face i-1/2
                dU[k] = (S)(-(p.rdx * (F[k] - Fl)));
            }
            if (HASB && p.bottom_source) dU[1] = (S)(dU[1] + bottom_term<P>(p, Blx, Brx, Lx[0], Rx[0], p.dx, p.rdx));
        }
        // ================= y direction (cell j, faces j-1/2 and j+1/2) =================
        {
            S Ly1[3], Ry1[3];
            recon_cell<P, RECON, false, false>(p, q0, q1, q2, bface_y(1), bface_y(2), bcell_of(1), Ly1, Ry1);   // cell j+1
            C G[3], s;
            central_upwind<P, false, 1, true>(p, Ry0, Ly1, G, s);           // face j+1/2
            if (STAGE == 1 && out_lane) msy = (S)mx((C)msy, s);
#pragma unroll
            for (int k = 0; k < 3; ++k) dU[k] = (S)(dU[k] - p.rdy * (G[k] - Gl[k]));
            if (HASB && p.bottom_source)
                dU[2] = (S)(dU[2] + bottom_term<P>(p, bface_y(0), bface_y(1), Ly0[0], Ry0[0], p.dy, p.rdy));
#pragma unroll
            for (int k = 0; k < 3; ++k) { Ly0[k] = Ly1[k]; Ry0[k] = Ry1[k]; Gl[k] = G[k]; }
        }
        // ================= per-cell sources (system.jl:71-73, list order) =================
#pragma unroll
        for (int pass = 0; pass < 2; ++pass) {
            const bool do_infil = (pass == 0) == (a.infil_first != 0);
            if (do_infil && a.infil_kind != SWE_B200_INFIL_NONE) {
                C f;
                if (a.factor) f = (C)(a.factor[(size_t)j * Nx + ic] * infil_t);
                else f = (C)infil_t;
                dU[0] = (S)(dU[0] - f);
            }
            if (!do_infil && a.rain_kind != SWE_B200_RAIN_NONE)
                dU[0] = (S)(dU[0] + (a.rainp ? (C)a.rainp[(size_t)j * Nx + ic] : rain));
        }
        // ================= store =================
        if (STAGE == 1) {
            if (out_lane) {
                S* o = a.out + (size_t)j * Nx + i;
                o[0] = dU[0];
                o[plane] = dU[1];
                o[2 * plane] = dU[2];
            }
        } else {
            // rungekutta.jl:39-49 stage 2, then friction (prev = U^1) and post-proc
            S o[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                o[k] = (S)(q0[k] + dt_s * dU[k]);
                o[k] = (S)((C)0.5 * (S)(un0[k] + o[k]));
            }
            const C Bm = bcell_of(0);
            if (p.friction != SWE_B200_FRICTION_NONE) {
                const C den = friction_denominator<P, false>(p, q0[0], q0[1], q0[2], Bm, dt_s);
                if constexpr (P::strict) {
                    o[1] = (S)(o[1] / den);
                    o[2] = (S)(o[2] / den);
                } else {
                    const C rden = rcp_fast(den);
                    o[1] = (S)(o[1] * rden);
                    o[2] = (S)(o[2] * rden);
                }
            }
            post_proc_cell<P>(p, o, Bm);
            if (out_lane) {
                S* w = a.out + (size_t)j * Nx + i;
                w[0] = o[0];
                w[plane] = o[1];
                w[2 * plane] = o[2];
            }
        }
        // ================= rotate =================
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            qm[k] = q0[k]; q0[k] = q1[k]; q1[k] = q2[k];
            un0[k] = un1[k]; un1[k] = un2[k];
        }
        ca[0] = ca[1]; cb[0] = cb[1]; ca[1] = ca[2]; cb[1] = cb[2]; ca[2] = ca[3]; cb[2] = cb[3];
    }

    if (STAGE == 1) {
        // CFL reduction: warp shuffle -> one atomicMax per warp and direction (non-negative bit patterns)
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

// ---- NeumannBC on the fused path (bc.jl:49-60) ----------------------------------------------------------------------
// The low-side ghosts are copies of interior cells (X: mirror, Y: image of the top interior rows) and are rebuilt /
// filled like the other BCs.  The high-side ghosts are not: the reference's ghost<->ghost copy (bc.jl:52-53, 57-58)
// leaves both layers holding the OLD outermost ghost cell, which therefore evolves on its own through the operators the
// reference applies to ALL cells -- sources (source_term_rain.jl:71-76, source_term_infiltration.jl:83-88), the stage
// update (rungekutta.jl:39-49), implicit friction (friction.jl:80-84), dry clean-up (timestepper.jl:41-51) -- with no
// flux and no bottom source (interior only).  Two small kernels per step carry that evolution:
//   k_neumann_dU     (before stage 2): dU := the per-cell sources on the outermost high ghost column / row, so that
//                    stage 2's on-the-fly rebuild U^1 = tail(U^n + dt dU) is right for the ghost source cells;
//   k_neumann_ghosts (before the tail): U^{n+1} of those cells (both RK2 stages, or the one ForwardEuler stage); the
//                    tail's ghost fill then copies them into the inner ghost layer.
template <class P>
__device__ __forceinline__ void neumann_line_cell(long long t, int Nx, int Ny, int& i, int& j) {
    // t in [0, Ny): column Nx-1, row t ; t in [Ny, Ny + Nx - 1): row Ny-1, column t - Ny
    if (t < Ny) { i = Nx - 1; j = (int)t; }
    else { i = (int)(t - Ny); j = Ny - 1; }
}
template <class P>
__global__ void k_neumann_dU(Par<typename P::C> p, typename P::S* D, const Ctl* ctl, const double* factor, const double* rainp,
                             int rain_kind, int infil_kind, int infil_first) {
    using S = typename P::S;
    using C = typename P::C;
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)p.Ny + (p.yhi_phys ? p.Nx - 1 : 0)) return;     // the ghost ROW only where the high y side is physical
    int i, j;
    neumann_line_cell<P>(t, p.Nx, p.Ny, i, j);
    const size_t q = (size_t)j * p.Nx + i;
    S d0 = 0;                                             // rungekutta.jl:28-30: out = 0, then the source terms in list order
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const bool do_infil = (pass == 0) == (infil_first != 0);
        if (do_infil && infil_kind != SWE_B200_INFIL_NONE) {
            const C f = factor ? (C)(factor[q] * ctl->infil) : (C)ctl->infil;
            d0 = (S)(d0 - f);
        }
        if (!do_infil && rain_kind != SWE_B200_RAIN_NONE) d0 = (S)(d0 + (rainp ? (C)rainp[q] : (C)ctl->rain));
    }
    D[q] = d0;
    D[q + p.plane] = 0;
    D[q + 2 * p.plane] = 0;
}
template <class P, bool RK2>
__global__ void k_neumann_ghosts(Par<typename P::C> p, Bathy<P> bt, const typename P::S* U, const typename P::S* D,
                                 typename P::S* out, const Ctl* ctl) {
    using S = typename P::S;
    using C = typename P::C;
    const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (t >= (long long)p.Ny + (p.yhi_phys ? p.Nx - 1 : 0)) return;
    int i, j;
    neumann_line_cell<P>(t, p.Nx, p.Ny, i, j);
    const size_t q = (size_t)j * p.Nx + i;
    S st[3], d[3], u1[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { st[k] = U[q + k * p.plane]; d[k] = D[q + k * p.plane]; }
    const C dt = (C)ctl->dt;
    const S dt_s = (S)ctl->dt_s;
    const C Bm = bt.cell(i, j);
    substep_tail<P>(p, st, d, dt, dt_s, Bm, u1);         // stage 1 (or the ForwardEuler step)
    if (!RK2) {
#pragma unroll
        for (int k = 0; k < 3; ++k) out[q + k * p.plane] = u1[k];
        return;
    }
    S o[3];                                               // stage 2 exactly as k_stage<.., STAGE = 2> finishes a cell
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        o[k] = (S)(u1[k] + dt_s * d[k]);
        o[k] = (S)((C)0.5 * (S)(st[k] + o[k]));
    }
    if (p.friction != SWE_B200_FRICTION_NONE) {
        const C den = friction_denominator<P, false>(p, u1[0], u1[1], u1[2], Bm, dt_s);
        if constexpr (P::strict) {
            o[1] = (S)(o[1] / den);
            o[2] = (S)(o[2] / den);
        } else {
            const C rden = rcp_fast(den);
            o[1] = (S)(o[1] * rden);
            o[2] = (S)(o[2] * rden);
        }
    }
    post_proc_cell<P>(p, o, Bm);
#pragma unroll
    for (int k = 0; k < 3; ++k) out[q + k * p.plane] = o[k];
}

// ForwardEuler tail, in place: vol = postproc(friction(U^n + dt*dU)) on the interior
// (forwardeuler.jl:42-44, friction.jl:80-84, timestepper.jl:41-51); ghosts follow in k_bc.
template <class P, bool HASB>
__global__ void k_update_fe(Par<typename P::C> p, Bathy<P> bt, const typename P::S* U, typename P::S* D, const Ctl* ctl) {
    using S = typename P::S;
    using C = typename P::C;
    const int i = blockIdx.x * blockDim.x + threadIdx.x + p.gc, j = blockIdx.y * blockDim.y + threadIdx.y + p.gc;
    if (i >= p.Nx - p.gc || j >= p.Ny - p.gc) return;
    const size_t q = (size_t)j * p.Nx + i;
    S st[3], d[3], o[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) { st[k] = U[q + k * p.plane]; d[k] = D[q + k * p.plane]; }
    substep_tail<P>(p, st, d, (C)ctl->dt, (S)ctl->dt_s, bt.cell(i, j), o);
#pragma unroll
    for (int k = 0; k < 3; ++k) D[q + k * p.plane] = o[k];
}

SWE_NS_END  // namespace swe

// kernels_unfused.cuh -- one CUDA kernel per reference operator (the per-function dispatch points
// of SURVEY 8b).  These reproduce the reference's unfused sequence including its buffer
// semantics (left/right face volumes, the shared wavespeed plane, pointwise kernels over ALL
// cells incl. ghosts).  In the strict policies every intermediate is bit-identical to the CPU
// oracle.  The fused stage kernels (kernels_fused.cuh) are the performance path.
#pragma once
#include <cstring>

#include "physics.cuh"

SWE_NS_BEGIN

// Device-resident control block: dt, t and the CFL reduction never visit the host inside a step
// (the reference syncs 4x per step: numericalflux.jl:32).
struct Ctl {
    unsigned long long maxspeed_bits[2];  // bit pattern of the non-negative realtype max per direction
    double dt;            // unrounded Float64 dt of stage 1 (rungekutta.jl:35)
    double dt_s;          // dt rounded to realtype (simulator.jl:151 current_timestep)
    double dt_report;     // current_timestep as the reference would report it: dt_s of the last real step
    double t;             // simulation time, realtype-rounded (simulator.jl:205)
    double t_end;         // simulate_to_time target
    double max_timestep;  // maximum_timestep (<=0: none)
    double max_dt_host;   // explicit max_dt of perform_step!
    double rain;          // current rain rate [m/s]  (source_term_rain.jl:30,45-54)
    double infil;         // current ConstantInfiltration rate or Horton time factor fc+(f0-fc)exp(-kt)
    long long nsteps;     // steps taken with t < t_end
    int use_t_end;        // 1: inside simulate_to_time (t_end applies), 0: max_dt_host (perform_step!)
    int match_endtime;
    unsigned int queue[2];  // work-queue heads of the two stage kernels of this step (kernels_fused_x2.cuh)
};

struct RainTable {
    int kind, n;
    double rate;
    double rates[16], times[16];
};
struct InfilPar {
    int kind;
    double cst, f0, fc, k;
};

template <class S> struct Bits;
template <> struct Bits<float> {
    __device__ static unsigned long long enc(float v) { return (unsigned long long)__float_as_uint(v); }
    __host__ __device__ static float dec(unsigned long long b) {
        unsigned u = (unsigned)b; float f;
#ifdef __CUDA_ARCH__
        f = __uint_as_float(u);
#else
        memcpy(&f, &u, 4);
#endif
        return f;
    }
};
template <> struct Bits<double> {
    __device__ static unsigned long long enc(double v) { return (unsigned long long)__double_as_longlong(v); }
    __host__ __device__ static double dec(unsigned long long b) {
        double f;
#ifdef __CUDA_ARCH__
        f = __longlong_as_double((long long)b);
#else
        memcpy(&f, &b, 8);
#endif
        return f;
    }
};

// block-level max of a non-negative value, then one atomicMax per block
template <class S>
__device__ __forceinline__ void block_atomic_max(S v, unsigned long long* dst) {
    __shared__ unsigned long long sm[32];
    unsigned long long b = Bits<S>::enc(v);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long t = __shfl_xor_sync(0xffffffffu, b, o);
        b = t > b ? t : b;
    }
    const int tid = threadIdx.x + blockDim.x * threadIdx.y;
    const int nw = (blockDim.x * blockDim.y + 31) >> 5;
    if ((tid & 31) == 0) sm[tid >> 5] = b;
    __syncthreads();
    if (tid < 32) {
        b = tid < nw ? sm[tid] : 0ull;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            unsigned long long t = __shfl_xor_sync(0xffffffffu, b, o);
            b = t > b ? t : b;
        }
        if (tid == 0 && b != 0ull) atomicMax(dst, b);
    }
}

// ---- control kernels ---------------------------------------------------------------------------------
// rain / infiltration rates at time t (start of the step, simulator.jl:159)
__device__ __forceinline__ void ctl_rates(Ctl* c, const RainTable& r, const InfilPar& f) {
    double rain = 0.0;
    if (r.kind == SWE_B200_RAIN_CONSTANT) rain = r.rate / 3600;
    else if (r.kind == SWE_B200_RAIN_TIME_DEPENDENT) {
        int index = 0;
        while (index < r.n - 1) {
            if (c->t < r.times[index + 1]) break;
            index += 1;
        }
        rain = r.rates[index] / 3600;
    }
    c->rain = rain;
    double inf = 0.0;
    if (f.kind == SWE_B200_INFIL_CONSTANT) inf = f.cst;
    else if (f.kind == SWE_B200_INFIL_HORTON) inf = f.fc + (f.f0 - f.fc) * exp(-f.k * c->t);
    c->infil = inf;
}

__device__ __forceinline__ void ctl_begin(Ctl* c, const RainTable& r, const InfilPar& f, double max_dt_host, int use_t_end,
                                          int match_endtime, double t_end, double max_timestep) {
    c->maxspeed_bits[0] = 0ull;
    c->maxspeed_bits[1] = 0ull;
    c->queue[0] = 0u;
    c->queue[1] = 0u;
    c->max_dt_host = max_dt_host;
    c->use_t_end = use_t_end;
    c->match_endtime = match_endtime;
    c->t_end = t_end;
    c->max_timestep = max_timestep;
    ctl_rates(c, r, f);
}
static __global__ void k_ctl_begin(Ctl* c, RainTable r, InfilPar f, double max_dt_host, int use_t_end, int match_endtime,
                                   double t_end, double max_timestep) {
    ctl_begin(c, r, f, max_dt_host, use_t_end, match_endtime, t_end, max_timestep);
}

// timestep_computer (simulator.jl:147-150): dt = min(cfl*min(dx/ax, dy/ay), max_dt), with
// max_dt = endtime - t (match_endtime) or Inf, capped by maximum_timestep (simulator.jl:200-203).
// Inside simulate_to_time (use_t_end) a step issued at t >= endtime gets dt = 0: an exact no-op,
// which is what lets the host batch steps without reading t back every step.
template <class S>
__global__ void k_ctl_dt(Ctl* c, double dx, double dy, S cfl) {
    const S ax = Bits<S>::dec(c->maxspeed_bits[0]);
    const S ay = Bits<S>::dec(c->maxspeed_bits[1]);
    double ddt = fmin(dx / ax, dy / ay);
    double max_dt = c->max_dt_host;
    if (c->use_t_end) max_dt = c->match_endtime ? (c->t_end - c->t) : (double)INFINITY;
    if (c->max_timestep > 0) max_dt = fmin(max_dt, c->max_timestep);
    double dt = fmin(cfl * ddt, max_dt);
    const bool past_end = c->use_t_end && !(c->t < c->t_end);
    if (past_end) dt = 0.0;
    c->dt = dt;
    c->dt_s = (double)(S)dt;
    if (!past_end) c->dt_report = c->dt_s;
}

// t += dt (simulator.jl:205), in realtype
template <class S>
__global__ void k_ctl_advance(Ctl* c) {
    if (c->use_t_end && !(c->t < c->t_end)) return;   // a step at t == t_end is an exact no-op
    S t = (S)c->t;
    t += (S)c->dt_s;
    c->t = (double)t;
    c->nsteps += 1;
}

// ---- pointwise kernels over ALL cells (for_each_cell, kernel_abstractions.jl:123-131) ------------------
template <class S>
__global__ void k_fill(S* a, long long n, S v) {
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) a[q] = v;
}

// out = state + dt*out (rungekutta.jl:39-41) ; stage 1 uses the unrounded dt, stage 2 the stored one
template <class P>
__global__ void k_axpy(typename P::S* out, const typename P::S* state, long long n, const Ctl* c, int use_rounded) {
    using S = typename P::S;
    using C = typename P::C;
    const C dt = (C)c->dt;
    const S dts = (S)c->dt_s;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        if (use_rounded) out[q] = (S)(state[q] + dts * out[q]);   // stage 2: dt is a realtype value
        else out[q] = (S)(state[q] + dt * out[q]);                // stage 1: the Float64 local
    }
}
// out = 0.5*(first + out) (rungekutta.jl:46-48)
template <class P>
__global__ void k_rkavg(typename P::S* out, const typename P::S* first, long long n) {
    using S = typename P::S;
    using C = typename P::C;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x)
        out[q] = (S)((C)0.5 * (S)(first[q] + out[q]));
}
// rain: out.h += r (source_term_rain.jl:73-75); infiltration: out.h -= f (source_term_infiltration.jl:85-87)
template <class P>
__global__ void k_point_source(typename P::S* out_h, long long n, const Ctl* c, int is_infil, const double* plane_values) {
    using S = typename P::S;
    using C = typename P::C;
    for (long long q = blockIdx.x * (long long)blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        if (is_infil) {
            C f = plane_values ? (C)(plane_values[q] * c->infil) : (C)c->infil;      // Horton factor plane
            out_h[q] = (S)(out_h[q] - f);
        } else {
            out_h[q] = (S)(out_h[q] + (plane_values ? (C)plane_values[q] : (C)c->rain));   // FunctionalRain values
        }
    }
}

template <class P, bool PURE>
__global__ void k_friction(Par<typename P::C> p, Bathy<P> bt, typename P::S* out, const typename P::S* prev,
                           const Ctl* c, double dt_host, int use_host_dt) {
    using S = typename P::S;
    using C = typename P::C;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= p.Nx || j >= p.Ny) return;
    const S dt_s = (S)(use_host_dt ? dt_host : c->dt_s);
    const size_t q = (size_t)j * p.Nx + i;
    const C Bm = PURE ? (C)0 : bt.cell(i, j);
    const C den = friction_denominator<P, PURE>(p, prev[q], prev[q + p.plane], prev[q + 2 * p.plane], Bm, dt_s);
    out[q] = (S)(out[q] / ((C)1 - dt_s * (C)0.0));
    out[q + p.plane] = (S)(out[q + p.plane] / den);
    out[q + 2 * p.plane] = (S)(out[q + 2 * p.plane] / den);
}

template <class P>
__global__ void k_postproc(Par<typename P::C> p, Bathy<P> bt, typename P::S* out) {
    using S = typename P::S;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= p.Nx || j >= p.Ny) return;
    const size_t q = (size_t)j * p.Nx + i;
    S o[3] = {out[q], out[q + p.plane], out[q + 2 * p.plane]};
    post_proc_cell<P>(p, o, bt.cell(i, j));
    out[q] = o[0];
    out[q + p.plane] = o[1];
    out[q + 2 * p.plane] = o[2];
}

// ---- reconstruct! : for_each_inner_cell(dir; ghostcells=1) / for_each_cell for NoReconstruction --------
template <class P, int RECON, bool PURE>
__global__ void k_reconstruct(Par<typename P::C> p, Bathy<P> bt, const typename P::S* in, typename P::S* left,
                              typename P::S* right, int dir) {
    using S = typename P::S;
    using C = typename P::C;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= p.Nx || j >= p.Ny) return;
    if (RECON == 1) {   // active index in [1, N-1), passive index interior only (grid.jl:213-244)
        if (dir == 0) { if (i < 1 || i >= p.Nx - 1 || j < p.gc || j >= p.Ny - p.gc) return; }
        else          { if (j < 1 || j >= p.Ny - 1 || i < p.gc || i >= p.Nx - p.gc) return; }
    }
    const size_t q = (size_t)j * p.Nx + i;
    const size_t st = dir == 0 ? 1 : (size_t)p.Nx;
    S l[3], c[3], r[3], L[3], R[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        c[k] = in[q + k * p.plane];
        l[k] = RECON ? in[q - st + k * p.plane] : c[k];
        r[k] = RECON ? in[q + st + k * p.plane] : c[k];
    }
    C Bl = 0, Br = 0, Bc = 0;
    if (!PURE) {
        if (RECON) { Bl = bt.face_left(i, j, dir); Br = bt.face_right(i, j, dir); }
        else Bc = bt.cell(i, j);
    }
    recon_cell<P, RECON, PURE>(p, l, c, r, Bl, Br, Bc, L, R);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        left[q + k * p.plane] = L[k];
        right[q + k * p.plane] = R[k];
    }
}

// ---- compute_flux! (numericalflux.jl:21-33): interior cells; also the per-direction max ----------------
template <class P, bool PURE, int DIR>
__global__ void k_flux(Par<typename P::C> p, typename P::S* out, const typename P::S* left, const typename P::S* right,
                       typename P::S* ws, Ctl* ctl) {
    using S = typename P::S;
    using C = typename P::C;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    S wsv = (S)0;
    if (i >= p.gc && i < p.Nx - p.gc && j >= p.gc && j < p.Ny - p.gc) {
        const size_t q = (size_t)j * p.Nx + i;
        const size_t st = DIR == 0 ? 1 : (size_t)p.Nx;
        S Rm[3], Lr[3], Rl[3], Lm[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            Rm[k] = right[q + k * p.plane];
            Lr[k] = left[q + st + k * p.plane];
            Rl[k] = right[q - st + k * p.plane];
            Lm[k] = left[q + k * p.plane];
        }
        C Fr[3], Fl[3], sr, sl;
        central_upwind<P, PURE, DIR>(p, Rm, Lr, Fr, sr);
        central_upwind<P, PURE, DIR>(p, Rl, Lm, Fl, sl);
        const C rd = DIR == 0 ? p.rdx : p.rdy;
#pragma unroll
        for (int k = 0; k < 3; ++k) out[q + k * p.plane] = (S)(out[q + k * p.plane] - rd * (Fr[k] - Fl[k]));
        wsv = (S)mx(sr, sl);
        ws[q] = wsv;
    }
    block_atomic_max<S>(wsv, &ctl->maxspeed_bits[DIR]);
}

// ---- evaluate_directional_source_term!(::SourceTermBottom) (source_term_bottom.jl:24-41) ---------------
template <class P>
__global__ void k_bottom(Par<typename P::C> p, Bathy<P> bt, typename P::S* out, const typename P::S* left,
                         const typename P::S* right, int dir) {
    using S = typename P::S;
    using C = typename P::C;
    const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y * blockDim.y + threadIdx.y;
    if (!(i >= p.gc && i < p.Nx - p.gc && j >= p.gc && j < p.Ny - p.gc)) return;
    const size_t q = (size_t)j * p.Nx + i;
    const C Br = bt.face_right(i, j, dir), Bl = bt.face_left(i, j, dir);
    const size_t m = q + (size_t)(dir == 0 ? 1 : 2) * p.plane;
    out[m] = (S)(out[m] + bottom_term<P>(p, Bl, Br, left[q], right[q], dir == 0 ? p.dx : p.dy, dir == 0 ? p.rdx : p.rdy));
}

// ---- update_bc! (bc.jl): for_each_ghost_cell(XDIR) then (YDIR); corners are never written ---------------
// One thread per (ghost index g, passive index); both sides handled by the same thread, in the
// reference's statement order (this matters for NeumannBC's ghost<->ghost copies, bc.jl:52-53).
template <class S>
__global__ void k_bc(S* d, int Nx, int Ny, long long plane, int gc, int bc, int dir, int lo_phys, int hi_phys) {
    const int N = dir == 0 ? Nx : Ny;
    const int np = dir == 0 ? Ny - 2 * gc : Nx - 2 * gc;
    const int pidx = blockIdx.x * blockDim.x + threadIdx.x;
    if (pidx >= np) return;
    const int pp = pidx + gc;
    const size_t sa = dir == 0 ? 1 : (size_t)Nx, sp = dir == 0 ? (size_t)Nx : 1;
    const int normal = dir == 0 ? 1 : 2;
    for (int g = 1; g <= gc; ++g) {
        for (int k = 0; k < 3; ++k) {
            S* base = d + (size_t)k * plane + (size_t)pp * sp;
            auto A = [&](int a1) -> S& { return base[(size_t)(a1 - 1) * sa]; };   // 1-based active index
            if (bc == SWE_B200_BC_PERIODIC) {
                if (lo_phys) A(g) = A(N + g - 2 * gc);
                if (hi_phys) A(N - (gc - g)) = A(gc + g);
            } else if (bc == SWE_B200_BC_NEUMANN) {
                if (lo_phys) A(g) = dir == 0 ? A(2 * gc - g + 1) : A(N + g - 2 * gc);
                if (hi_phys) A(N - (gc - g)) = A(N - g + 1);
            } else {
                const bool flip = k == normal;                 // a copy or a sign flip: exact for subnormals too
                if (lo_phys) { const S v = A(gc * 2 + 1 - g); A(g) = flip ? -v : v; }
                if (hi_phys) { const S v = A(N - gc - g + 1); A(N - gc + g) = flip ? -v : v; }
            }
        }
    }
}

// ---- device-side reductions for callbacks that only need scalars (examples/example_tools.jl:86-90,
// TotalWaterVolume: sum(collect(state.h) .- B_cell) over the interior; callbacks.jl:40-55): two passes, every
// partial in double and combined in a fixed order, so the result does not depend on the block schedule.
// slots: 0 sum(w - B_cell), 1 sum(hu), 2 sum(hv), 3 max(w - B_cell), 4 min(w - B_cell), 5 #wet cells
// (w - B_cell >= depth_cutoff), 6 max(|hu|, |hv|), 7 sum(w)
constexpr int DIAG_SLOTS = 8;
__device__ __forceinline__ void diag_combine(double a[DIAG_SLOTS], const double b[DIAG_SLOTS]) {
    a[0] += b[0]; a[1] += b[1]; a[2] += b[2];
    a[3] = fmax(a[3], b[3]); a[4] = fmin(a[4], b[4]);
    a[5] += b[5]; a[6] = fmax(a[6], b[6]); a[7] += b[7];
}
__device__ __forceinline__ void diag_identity(double a[DIAG_SLOTS]) {
    a[0] = a[1] = a[2] = a[5] = a[7] = 0.0;
    a[3] = -INFINITY; a[4] = INFINITY; a[6] = 0.0;
}
__device__ __forceinline__ void diag_block_reduce(double v[DIAG_SLOTS], double* out) {
    __shared__ double sh[8][DIAG_SLOTS];
    for (int o = 16; o > 0; o >>= 1) {
        double w[DIAG_SLOTS];
        for (int s = 0; s < DIAG_SLOTS; ++s) w[s] = __shfl_down_sync(0xffffffffu, v[s], o);
        diag_combine(v, w);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) for (int s = 0; s < DIAG_SLOTS; ++s) sh[warp][s] = v[s];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) diag_combine(v, sh[w]);
        for (int s = 0; s < DIAG_SLOTS; ++s) out[s] = v[s];
    }
}
template <class P>
__global__ void __launch_bounds__(256) k_diag_partial(const typename P::S* vol, Bathy<P> bt, int nx, int ny, int Nx,
                                                      int gc, long long plane, double cutoff, double* partial) {
    using S = typename P::S;
    double v[DIAG_SLOTS];
    diag_identity(v);
    const long long n = (long long)nx * ny;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(q % nx) + gc, j = (int)(q / nx) + gc;
        const size_t at = (size_t)j * Nx + i;
        const S w = vol[at], hu = vol[at + plane], hv = vol[at + 2 * plane];
        const S h = w - (S)bt.cell(i, j);            // in the storage type, like collect(state.h) .- bottom_avg
        double c[DIAG_SLOTS] = {(double)h, (double)hu, (double)hv, (double)h, (double)h, (double)h >= cutoff ? 1.0 : 0.0,
                                fmax(fabs((double)hu), fabs((double)hv)), (double)w};
        diag_combine(v, c);
    }
    diag_block_reduce(v, partial + (size_t)blockIdx.x * DIAG_SLOTS);
}
static __global__ void __launch_bounds__(256) k_diag_final(const double* partial, int nblocks, double* out) {
    double v[DIAG_SLOTS];
    diag_identity(v);
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) diag_combine(v, partial + (size_t)b * DIAG_SLOTS);
    diag_block_reduce(v, out);
}

// ---- per-cell comparison of two states on the same grid (swe_b200_compare_states): parity checks at sizes where
// copying both states to the host would dominate.  slots: 0..2 max|a-b| of w, hu, hv over the interior,
// 3..5 max|b| of w, hu, hv, 6 number of cells whose wet flag (w - B_cell >= depth_cutoff) differs, 7 / 8 wet cells of
// a / b, 9 number of cells with any NaN, 10..12 number of cells whose |a-b| of w, hu, hv exceeds thr[0..2] (absolute
// thresholds; 0 when thr is not given).
constexpr int CMP_SLOTS = 13;
struct CmpThr { double v[3]; int on; };
__device__ __forceinline__ void cmp_combine(double a[CMP_SLOTS], const double b[CMP_SLOTS]) {
    for (int s = 0; s < 6; ++s) a[s] = fmax(a[s], b[s]);
    for (int s = 6; s < CMP_SLOTS; ++s) a[s] += b[s];
}
__device__ __forceinline__ void cmp_block_reduce(double v[CMP_SLOTS], double* out) {
    __shared__ double sh[8][CMP_SLOTS];
    for (int o = 16; o > 0; o >>= 1) {
        double w[CMP_SLOTS];
        for (int s = 0; s < CMP_SLOTS; ++s) w[s] = __shfl_down_sync(0xffffffffu, v[s], o);
        cmp_combine(v, w);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) for (int s = 0; s < CMP_SLOTS; ++s) sh[warp][s] = v[s];
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < (int)(blockDim.x >> 5); ++w) cmp_combine(v, sh[w]);
        for (int s = 0; s < CMP_SLOTS; ++s) out[s] = v[s];
    }
}
template <class P>
__global__ void __launch_bounds__(256) k_cmp_partial(const typename P::S* A, const typename P::S* Bv, Bathy<P> bt, int nx,
                                                     int ny, int Nx, int gc, long long plane, double cutoff, CmpThr thr,
                                                     double* partial) {
    using S = typename P::S;
    double v[CMP_SLOTS];
    for (int s = 0; s < CMP_SLOTS; ++s) v[s] = 0.0;
    const long long n = (long long)nx * ny;
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n; q += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(q % nx) + gc, j = (int)(q / nx) + gc;
        const size_t at = (size_t)j * Nx + i;
        const S bc = (S)bt.cell(i, j);
        double c[CMP_SLOTS];
        bool nan = false;
        for (int k = 0; k < 3; ++k) {
            const double x = (double)A[at + k * plane], y = (double)Bv[at + k * plane];
            nan = nan || x != x || y != y;
            c[k] = fabs(x - y);
            c[3 + k] = fabs(y);
            c[10 + k] = thr.on && c[k] > thr.v[k] ? 1.0 : 0.0;
        }
        const bool wa = (double)(S)(A[at] - bc) >= cutoff, wb = (double)(S)(Bv[at] - bc) >= cutoff;
        c[6] = wa != wb ? 1.0 : 0.0; c[7] = wa ? 1.0 : 0.0; c[8] = wb ? 1.0 : 0.0; c[9] = nan ? 1.0 : 0.0;
        cmp_combine(v, c);
    }
    cmp_block_reduce(v, partial + (size_t)blockIdx.x * CMP_SLOTS);
}
static __global__ void __launch_bounds__(256) k_cmp_final(const double* partial, int nblocks, double* out) {
    double v[CMP_SLOTS];
    for (int s = 0; s < CMP_SLOTS; ++s) v[s] = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) cmp_combine(v, partial + (size_t)b * CMP_SLOTS);
    cmp_block_reduce(v, out);
}

// ---- iteration spaces of the backend kernel surface (kernel_abstractions.jl:66-131, grid.jl:213-253) ------
// kind 0: for_each_cell, 1: for_each_inner_cell(dir; ghostcells), 2: for_each_ghost_cell(dir).
// Each work-item records the 1-based indices it would hand to the closure.
static __global__ void k_enumerate(int kind, int dir, int g, int Nx, int Ny, int gc, long long* out) {
    const int a = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y * blockDim.y + threadIdx.y;
    if (kind == 0) {
        if (a >= Nx || b >= Ny) return;
        long long* o = out + 2 * ((long long)b * Nx + a);
        o[0] = a + 1; o[1] = b + 1;
    } else if (kind == 1) {
        // ndrange = inner_cells(grid, dir, g): active extent N_d - 2g, passive extent N - 2gc
        const int na = dir == 0 ? Nx - 2 * g : Nx - 2 * gc;
        const int nb = dir == 0 ? Ny - 2 * gc : Ny - 2 * g;
        if (a >= na || b >= nb) return;
        const int I1 = a + 1, I2 = b + 1;   // @index(Global, Cartesian)
        long long* o = out + 6 * ((long long)b * na + a);
        if (dir == 0) {   // left/middle/right_cell(grid, I, XDIR, g) grid.jl:223-240
            o[0] = I1 + g - 1; o[1] = I2 + gc; o[2] = I1 + g; o[3] = I2 + gc; o[4] = I1 + g + 1; o[5] = I2 + gc;
        } else {
            o[0] = I1 + gc; o[1] = I2 + g - 1; o[2] = I1 + gc; o[3] = I2 + g; o[4] = I1 + gc; o[5] = I2 + g + 1;
        }
    } else {
        // ndrange = ghost_cells(grid, dir); middle_cell(grid, I, dir, 0)
        const int na = dir == 0 ? gc : Nx - 2 * gc;
        const int nb = dir == 0 ? Ny - 2 * gc : gc;
        if (a >= na || b >= nb) return;
        long long* o = out + 2 * ((long long)b * na + a);
        if (dir == 0) { o[0] = a + 1; o[1] = b + 1 + gc; }
        else          { o[0] = a + 1 + gc; o[1] = b + 1; }
    }
}

SWE_NS_END  // namespace swe

// kernels_fused_d1.cuh -- the Float64 performance path: the fused RK2 stage kernels of kernels_fused_x2.cuh rebuilt
// for ONE Float64 column per thread (Float64 is the reference's own GPU precision: make_cuda_backend() has no
// realtype argument, backends/kernel_abstractions.jl:36).
//
// Why a kernel of its own: the generic k_stage<P> (kernels_fused.cuh) executes ~700 instructions per cell and stage in
// Float64 (profiles/r01j: 15 % of the HBM roofline at 1024^2, 33 % at 4096^2) -- selects, 64-bit register moves of the
// rotating windows and IEEE division / square-root slow paths, not the DP pipe.  Everything that made the packed
// Float32 kernel fast and is not packing carries over, because a double is exactly as wide as the float2 pair:
//   * the per-thread cp.async (LDGSTS) rings in shared memory, rows j+3 / j+4 in flight, no barrier in the kernel;
//   * the y march unrolled by 3 = the period of the register window (no rotation moves), FAST / generic trips;
//   * persistent warps pulling (tile, rows) items from one atomic counter, guided schedule (EngineT::build_items);
//   * dry face sides masked on the INPUT (h*m, (1/h*)*m), -a- carried, slope clamp as (s > s1 ? s1 : max(s, s2)),
//     desingularisation as h + relu(k - h)^2 / 2k, one-reciprocal friction factors, one dry-skip vote per row;
//   * reciprocals and square roots from the hardware seeds + Newton steps (no slow-path calls);
//   * minmod with three compare-and-selects: sign(f) min(theta min(|f|,|b|), (|f|+|b|)/2) for f, b of equal sign
//     (bitwise the value of minmod(theta f, (f+b)/2, theta b), linearreconstruction.jl:28-43).
// A warp covers 32 columns and produces the inner 28; x-neighbours come by warp shuffles.
// Semantics are those of kernels_fused.cuh (same reference citations); parity: tests/test_gpu_parity.py compares this
// path against the CPU oracle at rel-Linf <= 1e-12 (the north_star Float64 tolerance).
#pragma once
#include <cstddef>
#include <type_traits>

#include "kernels_fused.cuh"

SWE_NS_BEGIN
namespace d1 {

typedef double R;

constexpr int VW = 28;       // valid (output) columns per warp: lanes 2..29
constexpr int WPB = 4;       // warps per block
constexpr int NT = WPB * 32;

#ifndef D1_MINB
#define D1_MINB 4            // resident blocks per SM, stage 1 (128 registers)
#endif
#ifndef D1_MINB2
#define D1_MINB2 3           // stage 2 (168 registers)
#endif
template <int STAGE>
constexpr int minb() { return STAGE == 2 ? D1_MINB2 : D1_MINB; }

__device__ __forceinline__ R vmin(R a, R b) { return a < b ? a : b; }
__device__ __forceinline__ R vmax(R a, R b) { return a > b ? a : b; }
// The FP64 pipe is the top pipe (profiles/r02f_ncu_d1_f64_4096.txt: 62 % busy; DSETP runs there too), but the kernel is issue-bound:
// moving 64-bit comparisons to the integer ALU (two ISETP on the bit patterns of non-negative operands) relieves the pipe and
// costs instructions -- measured slower (D1_HYBRID below).  What stays on the ALU are the tests that are a single integer
// instruction: the sign bit (relu) and the exponent word (the normal-number test of sqrt_n).
__device__ __forceinline__ R vmin_nn(R a, R b) { return __double_as_longlong(a) < __double_as_longlong(b) ? a : b; }
#ifndef D1_HYBRID
#define D1_HYBRID 3
#endif
#if D1_HYBRID >= 3
__device__ __forceinline__ R vmax_nn(R a, R b) { return a > b ? a : b; }
#else
__device__ __forceinline__ R vmax_nn(R a, R b) { return __double_as_longlong(a) > __double_as_longlong(b) ? a : b; }
#endif
__device__ __forceinline__ R relu(R x) { return __double2hiint(x) < 0 ? 0.0 : x; }
__device__ __forceinline__ R rcp_n(R x) {                // hardware seed (~2^-20) + two Newton steps: ~1 ulp
    R r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    R e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    return fma(r, e, r);
}
__device__ __forceinline__ R sqrt_n(R x) {               // x >= 0: rsqrt seed (~2^-22) + one Newton step (~2^-44) on
    R y;                                                 // 1/sqrt(x), then one correction step on sqrt(x) itself (~1 ulp)
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = y * fma(-0.5 * x * y, y, 1.5);
    R s = x * y;
    s = fma(fma(-s, s, x), 0.5 * y, s);
    return __double2hiint(x) >= 0x00100000 ? s : 0.0;    // normal x only (rsqrt(0) = inf; the flush-to-zero seed does not see subnormals)
}
__device__ __forceinline__ R shfl_up(R v) { return __shfl_up_sync(0xffffffffu, v, 1); }
__device__ __forceinline__ R shfl_dn(R v) { return __shfl_down_sync(0xffffffffu, v, 1); }

// kernel parameters: the Float64 scalars of the reference (constant bank)
struct Par1 {
    int Nx, Ny;
    long long plane;
    int Bp;                       // pitch of the corner array
    int bc, friction, bottom_source, ylo_phys, yhi_phys;
    int rain_kind, infil_kind, infil_first;
    int nitems;                   // work items of one launch (Args1::items)
    R cutoff, kappa, constB;
    R theta, g, K /*0.5 rho g*/, i2k /*1/2k*/, rho, nrdx, nrdy, botx, boty, Cz;
};

struct Args1 {
    Par1 p;
    const R* U;
    const R* D;
    R* out;
    const R* B;
    const R* wrapLo;              // corner rows of the wrap-around image's source rows (see StageArgs), or nullptr
    const R* wrapHi;
    const double* factor;
    Ctl* ctl;
    const int4* items;            // (tile, first interior row, number of rows (multiple of 3 except at the top), 0)
};

// ---- minmod-limited slope from the backward / forward differences (linearreconstruction.jl:28-43) ----------------
__device__ __forceinline__ R slope(R theta, R b, R f) {
    const R af = fabs(f), ab = fabs(b);
    // 64-bit comparisons: one DSETP on the FP64 pipe, or two ISETP on the integer ALU (non-negative operands).  D1_HYBRID: 0 =
    // both minima on the ALU, 1 = the first on the FP64 pipe, 2 = both (sign / zero tests stay on the ALU), 3 = every 64-bit
    // comparison on the FP64 pipe.  Measured (profiles/r02v_ab_d1_hybrid.txt): config 2 at 4096^2 0.696 / 0.684 / 0.666 ms/step
    // for 0 / 1 / 2, 0.664 for 3 (the default) -- the kernel is issue-bound, so the fewer instructions win although the FP64 pipe is
    // the busier one; only the tests that are ONE integer instruction (sign bit, exponent word) stay on the ALU.
#if D1_HYBRID >= 1
    const R m1 = vmin(af, ab);
#else
    const R m1 = vmin_nn(af, ab);
#endif
#if D1_HYBRID >= 2
    const R m2 = vmin(theta * m1, 0.5 * (af + ab));
#else
    const R m2 = vmin_nn(theta * m1, 0.5 * (af + ab));
#endif
    // equal strict signs, tested on the sign bits (a product f * b would underflow for tiny differences)
#if D1_HYBRID >= 3
    const bool same = ((__double2hiint(f) ^ __double2hiint(b)) >= 0) && m1 > 0.0;
#else
    const bool same = ((__double2hiint(f) ^ __double2hiint(b)) >= 0) && __double_as_longlong(m1) != 0ll;
#endif
    return same ? copysign(m2, f) : 0.0;
}

// ---- face values of one cell from its centre values and limited slopes (linearreconstruction.jl:73-100;
// noreconstruction.jl:30-46).  tBl / tBr = 2 * face bathymetry (sum of the two corners). -------------------------
template <int RECON>
__device__ __forceinline__ void recon1(const Par1& p, const R c[3], const R bwd[3], const R fwd[3], R tBl, R tBr, R Bc,
                                       R L[3], R Rr[3]) {
    if (RECON == 0) {
        L[0] = Rr[0] = c[0] - Bc;
        L[1] = Rr[1] = c[1];
        L[2] = Rr[2] = c[2];
        return;
    }
    R s[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) s[k] = slope(p.theta, bwd[k], fwd[k]);
    // slope clamp of w (no negative face depth): (w - s/2 < Bl ? 2(w-Bl) : w + s/2 < Br ? 2(Br-w) : s)
    const R w2 = c[0] + c[0];
    const R s1 = w2 - tBl, s2 = tBr - w2;
    s[0] = s[0] > s1 ? s1 : vmax(s[0], s2);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        L[k] = fma(s[k], -0.5, c[k]);
        Rr[k] = fma(s[k], 0.5, c[k]);
    }
    L[0] = fma(tBl, -0.5, L[0]);
    Rr[0] = fma(tBr, -0.5, Rr[0]);
}

// desingularised depth of a NON-NEGATIVE depth (equation.jl:56-62): max(h, min(h^2/2k + k/2, k)) = h + relu(k - h)^2 / 2k
__device__ __forceinline__ R desing_pos(const Par1& p, R h) {
    const R rl = relu(p.kappa - h);
    return fma(rl * rl, p.i2k, h);
}

// ---- one side of a face: physical flux, u+c and c-u with the dry gate folded into the inputs
// (swe_2D.jl:43-81, equation.jl:40-62, centralupwind.jl:54-63) ----------------------------------------------------
template <int DIR>
__device__ __forceinline__ void side1(const Par1& p, R h, R hu, R hv, R F[3], R& ep, R& en) {
#ifndef D1_MASKMUL
#define D1_MASKMUL 0         // 1: hm = h * m (one DMUL) instead of the select (two SEL)
#endif
    const bool wet = h > p.cutoff;
    const R m = wet ? 1.0 : 0.0;
#if D1_MASKMUL
    const R hm = h * m;
#else
    const R hm = wet ? h : 0.0;
#endif
    const R pres = p.K * (hm * hm);
    const R c = sqrt_n(p.g * hm);
    const R rm = rcp_n(desing_pos(p, hm)) * m;                     // (a multiply, not a select: no branch around the Newton steps)
    const R u = hu * rm, v = hv * rm;
    const R rh = p.rho * hm;
    R un;
    if (DIR == 0) {
        F[0] = rh * u;
        F[1] = fma(F[0], u, pres);
        F[2] = F[0] * v;
        un = u;
    } else {
        F[0] = rh * v;
        F[1] = F[0] * u;
        F[2] = fma(F[0], v, pres);
        un = v;
    }
    ep = un + c;
    en = c - un;
}

// ---- Kurganov-Petrova central-upwind flux through one face (centralupwind.jl:48-77) ------------------------------
template <int DIR, bool SPEED>
__device__ __forceinline__ void cu1(const Par1& p, const R Um[3], const R Up[3], R F[3], R& speed) {
    R Fm[3], Fp[3], epm, enm, epp, enp;
    side1<DIR>(p, Um[0], Um[1], Um[2], Fm, epm, enm);
    side1<DIR>(p, Up[0], Up[1], Up[2], Fp, epp, enp);
    const R ap = relu(vmax(epm, epp));                   //  a+
    const R an = relu(vmax(enm, enp));                   // -a-
    const R d = ap + an;
    const bool ok = !(d < p.kappa);                      // :67-69
    const R r = ok ? rcp_n(d) : 0.0;
    const R wp = ap * r, wn = an * r;
    const R c2 = ap * wn;
#pragma unroll
    for (int k = 0; k < 3; ++k) F[k] = fma(c2, Um[k] - Up[k], fma(wn, Fp[k], wp * Fm[k]));
    speed = (SPEED && ok) ? vmax_nn(ap, an) : 0.0;
}

// ---- pointwise tail of a substep (rungekutta.jl:39-41, friction.jl:62-84, timestepper.jl:41-51): returns the factor
// the momenta are multiplied by -- 1/(1 - dt*c) of the implicit friction times the wet flag of the dry clean-up --
// and raises o0 to the cell bathymetry.
template <int FRIC>
__device__ __forceinline__ R tail_factor1(const Par1& p, const R prev[3], R Bc, R czdt, R& o0) {
    R fac = (o0 - Bc) < p.cutoff ? 0.0 : 1.0;
    o0 = vmax(o0, Bc);
    if (FRIC != 0) {
        const R hsa = desing_pos(p, fabs(prev[0] - Bc));     // |h*|; h* = copysign(|h*|, h)
        const R sm = sqrt_n(fma(prev[1], prev[1], prev[2] * prev[2]));   // |(hu, hv)| = speed * |h*|
        // 1 / (1 + dt Cz speed / law(h*)) with speed = |m| / h*  =  law h* / (law h* + dt Cz |m|): one reciprocal
        R lh;                                                 // law(h*) * h*
        if (FRIC == 1) lh = hsa * hsa * hsa;                  // h*^2            (friction_bh2021)
        else {
            const R cb = cbrt(hsa);
            lh = cb * hsa * hsa;                              // cbrt(h*) h*     (friction_bsa2012)
            if (p.friction != SWE_B200_FRICTION_BSA2012) lh *= hsa;   // cbrt(h*) h*^2 (friction_fcg2016)
        }
        fac *= lh * rcp_n(fma(sm, czdt, lh));
    }
    return fac;
}

__device__ __forceinline__ void cp_async8(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_1() { asm volatile("cp.async.wait_group 1;" ::: "memory"); }

// Per-thread rings in shared memory, exactly the layout of the Float32 kernel (a double is as wide as its float2):
//   U : U^n rows (stage 2 reads row j again for the Heun average -> 2 banks x 3 slots; stage 1: 3 slots)
//   D : dU rows (stage 2)                                         -> 3 slots
//   E / E2 : the two bathymetry corners B[ic], B[ic+1] of a corner row -> 2 banks x 3 slots each
template <int STAGE>
struct Rings {
    R U[STAGE == 2 ? 2 : 1][3][3][NT];      // [bank][slot][variable][thread]
    R D[STAGE == 2 ? 3 : 1][3][STAGE == 2 ? NT : 1];
    R E[2][3][NT];
    R E2[2][3][NT];
};

template <int RECON, bool HASB, int STAGE, int FRIC>
__global__ void __launch_bounds__(NT, minb<STAGE>()) k_stage_d1(const Args1 a) {
    __shared__ Rings<STAGE> sm;
    constexpr int UB = STAGE == 2 ? 2 : 1;               // banks of the U ring
    const Par1& p = a.p;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int Nx = p.Nx, Ny = p.Ny;
    R msx = 0.0, msy = 0.0;                              // running CFL maxima of this thread (stage 1)
    const int total_items = p.nitems;
  for (;;) {
    int item = 0;
    if (lane == 0) item = (int)atomicAdd(&a.ctl->queue[STAGE - 1], 1u);
    item = __shfl_sync(0xffffffffu, item, 0);
    if (item >= total_items) break;
    const int4 it = a.items[item];
    const int tile = it.x;
    const int i = tile * VW + lane;                      // this lane's column
    const int ic = min(i, Nx - 1);                       // clamped for loads
    const int ja = it.y;
    const int jb = min(ja + it.z, Ny - 2);
    const bool out_lane = lane >= 2 && lane <= 29 && i < Nx - 2;
    const bool xface_ok = lane >= 1 && lane <= 29 && i >= 1 && i <= Nx - 3;   // face (i | i+1) touches an interior cell
    const long long plane = p.plane;
    const int Bp = p.Bp;

    R dt = 0.0, czdt = 0.0;
    int xs = ic;
    bool xflip = false;
    if (STAGE == 2) {
        dt = a.ctl->dt_s;
        czdt = p.Cz * dt;
        xs = bc_source(ic, Nx, 2, p.bc, true, true, xflip);   // a ghost column: where the BC copies it from
    }
    const bool xghost = xs != ic;
    // per-cell sources (system.jl:71-73): rain - infiltration is one constant per step unless a Horton factor plane exists
    const R rain = p.rain_kind != SWE_B200_RAIN_NONE ? a.ctl->rain : 0.0;
    const R infil_t = p.infil_kind != SWE_B200_INFIL_NONE ? a.ctl->infil : 0.0;
    const bool has_factor = a.factor != nullptr;
    const R src = has_factor ? rain : rain - infil_t;

    struct Corner { R a, b; };                           // B[ic, row], B[ic+1, row]
    auto ty_of = [&](const Corner& c) -> R { return HASB ? c.a + c.b : 2.0 * p.constB; };   // 2 * B_face_y above that corner row
    auto bcell = [&](R tylo, R tyhi) -> R { return HASB ? 0.25 * (tylo + tyhi) : p.constB; };
    auto row_source = [&](int j, bool& yflip) -> int {   // stage 2: the row the y BC copies row j from
        yflip = false;
        if (STAGE == 1) return j;
        return bc_source(j, Ny, 2, p.bc, p.ylo_phys, p.yhi_phys, yflip, 1);
    };
    // stage 1: q = U^n ; stage 2: q = U^1 = BC(postproc(friction(U^n + dt*dU))) rebuilt on the fly
    auto finish_row = [&](bool remapped, int ys, bool yflip, const R u[3], const R d[3], R Bc_here, R q[3]) {
        if (STAGE == 1) {
#pragma unroll
            for (int k = 0; k < 3; ++k) q[k] = u[k];
        } else {
            R Bm = Bc_here;
            const R* wr = !HASB ? nullptr : (ys < 2 ? a.wrapLo : (ys >= Ny - 2 ? a.wrapHi : nullptr));   // (not remapped there: ys = row)
            if (HASB && (xghost || remapped || wr)) {         // ghost: bathymetry of the source cell
                const R* r0 = wr ? wr + (size_t)(ys < 2 ? ys : ys - (Ny - 2)) * Bp + xs : a.B + (size_t)ys * Bp + xs;
                Bm = 0.25 * (r0[0] + r0[1] + r0[Bp] + r0[Bp + 1]);
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) q[k] = fma(dt, d[k], u[k]);
            const R fac = tail_factor1<FRIC>(p, u, Bm, czdt, q[0]);
            q[1] *= fac;
            q[2] *= fac;
            if (xflip) q[1] = -q[1];
            if (yflip) q[2] = -q[2];
        }
    };

    // ---- fetch of state row r / corner row r into the rings; running byte pointers advance by one row per fetch ----
    const long long rowB = (long long)Nx * 8, planeB = plane * 8;
    const long long dOff = (const char*)a.D - (const char*)a.U;
    const char* pf = (const char*)(a.U + (size_t)(ja + 2) * Nx + xs);     // next state row to fetch
    const char* pc = (const char*)(a.B + (size_t)(ja + 3) * Bp + ic);     // next corner row to fetch
    char* ps = (char*)(a.out + (size_t)ja * Nx + i);                      // row j of the output
    const double* pfac = a.factor + (size_t)ja * Nx + ic;                 // Horton factor plane, row j
    int ys_next = 0, ys_q = 0;                            // BC source row of the state rows in flight (generic path)
    bool yflip_next = false, remap_next = false, yflip_q = false, remap_q = false;
    char* const smb = reinterpret_cast<char*>(&sm);
    constexpr unsigned U_BANK = 3 * 3 * NT * 8, U_SLOT = 3 * NT * 8, U_VAR = NT * 8;
    constexpr unsigned D_OFF = offsetof(Rings<STAGE>, D), E_OFF = offsetof(Rings<STAGE>, E), E2_OFF = offsetof(Rings<STAGE>, E2);
    constexpr unsigned E_BANK = 3 * NT * 8, E_SLOT = NT * 8;
    char* const smU = smb + tid * 8;
    char* const smD = smb + D_OFF + tid * 8;
    char* const smE = smb + E_OFF + tid * 8;
    char* const smE2 = smb + E2_OFF + tid * 8;
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
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            cp_async8(smU + ubank + slot * U_SLOT + k * U_VAR, g + k * planeB);
            if (STAGE == 2) cp_async8(smD + slot * U_SLOT + k * U_VAR, g + dOff + k * planeB);
        }
    };
    auto fetch_corner = [&](auto fast_tag, int row, unsigned ebank, int slot) {
        constexpr bool FAST = decltype(fast_tag)::value;
        if (HASB) {
            const char* g = pc;
            if (!FAST && row > Ny) g = (const char*)(a.B + (size_t)Ny * Bp + ic);
            pc += (long long)Bp * 8;
            cp_async8(smE + ebank + slot * E_SLOT, g);
            cp_async8(smE2 + ebank + slot * E_SLOT, g + 8);
        }
    };
    fetch_row(std::false_type{}, ja + 2, 0u, 2);
    fetch_corner(std::false_type{}, ja + 3, E_BANK, 0);
    cp_async_commit();
    fetch_row(std::false_type{}, ja + 3, UB == 2 ? U_BANK : 0u, 0);
    fetch_corner(std::false_type{}, ja + 4, E_BANK, 1);
    cp_async_commit();

    // ---- prologue: state rows ja-2 .. ja+1 and corner rows ja-2 .. ja+2 by direct loads ----
    R q[3][3];                                            // q[(j - ja) % 3] = row j of the (rebuilt) state
    R dprev[3], Ry0[3], Gl[3], hsy;                       // y state of cell j: q(j+1)-q(j), upper face values, lower flux, hL+hR
    Corner Ej;                                            // corner row j
    R tyj, ty2;                                           // 2*B_face_y above corner rows j and j+2
    {
        Corner c[5];
#pragma unroll
        for (int r = 0; r < 5; ++r) {
            c[r].a = c[r].b = 0.0;
            if (HASB) {
                const R* g = a.B + (size_t)(ja - 2 + r) * Bp + ic;
                c[r].a = g[0];
                c[r].b = g[1];
            }
        }
        R u[4][3], d[4][3];
        int ys[4];
        bool yf[4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            ys[r] = row_source(ja - 2 + r, yf[r]);
            const size_t off = (size_t)ys[r] * Nx + xs;
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                u[r][k] = a.U[off + k * plane];
                d[r][k] = STAGE == 2 ? a.D[off + k * plane] : 0.0;
            }
        }
        if (STAGE == 2) {                                 // rows ja, ja+1 of U^n are read again by the Heun average
#pragma unroll
            for (int k = 0; k < 3; ++k) { sm.U[0][0][k][tid] = u[2][k]; sm.U[0][1][k][tid] = u[3][k]; }
        }
        if (HASB) {                                       // corner rows ja+1, ja+2 are read again for the x faces
            sm.E[0][1][tid] = c[3].a; sm.E2[0][1][tid] = c[3].b;
            sm.E[0][2][tid] = c[4].a; sm.E2[0][2][tid] = c[4].b;
        }
        const R tm2 = ty_of(c[0]), tm1 = ty_of(c[1]), t0 = ty_of(c[2]), t1 = ty_of(c[3]), t2 = ty_of(c[4]);
        R qa[3], qb[3];
        finish_row(ys[0] != ja - 2, ys[0], yf[0], u[0], d[0], bcell(tm2, tm1), qa);
        finish_row(ys[1] != ja - 1, ys[1], yf[1], u[1], d[1], bcell(tm1, t0), qb);
        finish_row(ys[2] != ja, ys[2], yf[2], u[2], d[2], bcell(t0, t1), q[0]);
        finish_row(ys[3] != ja + 1, ys[3], yf[3], u[3], d[3], bcell(t1, t2), q[1]);
        R b0[3], b1[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            b0[k] = qb[k] - qa[k]; b1[k] = q[0][k] - qb[k]; dprev[k] = q[1][k] - q[0][k];
        }
        R Lt[3], Rp[3], Ly0[3];
        recon1<RECON>(p, qb, b0, b1, tm1, t0, bcell(tm1, t0), Lt, Rp);              // cell ja-1: upper face only
        recon1<RECON>(p, q[0], b1, dprev, t0, t1, bcell(t0, t1), Ly0, Ry0);         // cell ja
        R s = 0.0;
        Gl[0] = Gl[1] = Gl[2] = 0.0;
        if (__any_sync(0xffffffffu, vmax(Rp[0], Ly0[0]) > p.cutoff)) cu1<1, STAGE == 1>(p, Rp, Ly0, Gl, s);   // face ja-1/2
        if (STAGE == 1 && out_lane) msy = vmax_nn(msy, s);
        hsy = Ly0[0] + Ry0[0];
        Ej = c[2];
        tyj = t0;
        ty2 = t2;
    }

    // ---- march: 3 rows per trip; rows beyond jb (last chunk only) are computed on clamped data and not stored ----
    int bank = 0;                                         // (j0 - ja) / 3 mod 2
    auto trip = [&](auto fast_tag, const int j0) {
        constexpr bool FAST = decltype(fast_tag)::value;
        const unsigned uU[2] = {UB == 2 && bank ? U_BANK : 0u, UB == 2 && !bank ? U_BANK : 0u};
        const unsigned uE[2] = {bank ? E_BANK : 0u, bank ? 0u : E_BANK};
#pragma unroll
        for (int Rr = 0; Rr < 3; ++Rr) {
            const int j = j0 + Rr;
            R* q0 = q[Rr];
            R* q1 = q[(Rr + 1) % 3];
            R* q2 = q[(Rr + 2) % 3];
            // ---- arrivals: state row j+2 and corner row j+3 (rows j+3 / j+4 stay in flight) ----
            cp_async_wait_1();
            const bool yflip2 = !FAST && yflip_q, remap2 = !FAST && remap_q;
            const int ys2 = remap2 ? ys_q : j + 2;
            R ru[3], rd[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                ru[k] = *reinterpret_cast<const R*>(smU + uU[((Rr + 2) / 3) & 1] + ((Rr + 2) % 3) * U_SLOT + k * U_VAR);
                rd[k] = STAGE == 2 ? *reinterpret_cast<const R*>(smD + ((Rr + 2) % 3) * U_SLOT + k * U_VAR) : 0.0;
            }
            Corner E1, E3;                                // corner rows j+1 and j+3
            E1.a = E1.b = 0.0; E3 = E1;
            if (HASB) {
                E1.a = *reinterpret_cast<const R*>(smE + uE[((Rr + 1) / 3) & 1] + ((Rr + 1) % 3) * E_SLOT);
                E1.b = *reinterpret_cast<const R*>(smE2 + uE[((Rr + 1) / 3) & 1] + ((Rr + 1) % 3) * E_SLOT);
                E3.a = *reinterpret_cast<const R*>(smE + uE[1] + Rr * E_SLOT);
                E3.b = *reinterpret_cast<const R*>(smE2 + uE[1] + Rr * E_SLOT);
            }
            fetch_row(fast_tag, j + 4, uU[((Rr + 4) / 3) & 1], (Rr + 1) % 3);    // two rows ahead of the arrivals
            fetch_corner(fast_tag, j + 5, uE[((Rr + 5) / 3) & 1], (Rr + 2) % 3);
            cp_async_commit();
            const R ty1 = ty_of(E1), ty3 = ty_of(E3);
            finish_row(remap2, ys2, yflip2, ru, rd, bcell(ty2, ty3), q2);

            // ================= reconstruction: row j in x, cell j+1 in y =================
            R Lx[3], Rx[3], Ln[3], tBl, tBr, fwd[3], Ly1[3], Ry1[3];
            {
                R bx[3], fx[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    bx[k] = q0[k] - shfl_up(q0[k]);           // q(i) - q(i-1)
                    fx[k] = shfl_dn(q0[k]) - q0[k];           // q(i+1) - q(i)
                }
                if (HASB) {
                    tBl = Ej.a + E1.a;
                    tBr = Ej.b + E1.b;
                } else tBl = tBr = 2.0 * p.constB;
                recon1<RECON>(p, q0, bx, fx, tBl, tBr, bcell(tyj, ty1), Lx, Rx);
#pragma unroll
                for (int k = 0; k < 3; ++k) Ln[k] = shfl_dn(Lx[k]);   // left face values of cell i+1
#pragma unroll
                for (int k = 0; k < 3; ++k) fwd[k] = q2[k] - q1[k];
                recon1<RECON>(p, q1, dprev, fwd, ty1, ty2, bcell(ty1, ty2), Ly1, Ry1);   // cell j+1
            }
            // ================= fluxes: face (i|i+1) of row j and face j+1/2 =================
            // one warp-uniform decision per row: a warp whose x and y faces are all dry skips both evaluations
            R F[3], G[3];
            const bool wet = vmax(vmax(Rx[0], Ln[0]), vmax(Ry0[0], Ly1[0])) > p.cutoff;
            if (__any_sync(0xffffffffu, wet)) {
                R sx, sy;
                cu1<0, STAGE == 1>(p, Rx, Ln, F, sx);
                cu1<1, STAGE == 1>(p, Ry0, Ly1, G, sy);
                if (STAGE == 1 && (FAST || j < jb)) {
                    if (xface_ok) msx = vmax_nn(msx, sx);
                    if (out_lane) msy = vmax_nn(msy, sy);
                }
            } else {
#pragma unroll
                for (int k = 0; k < 3; ++k) F[k] = G[k] = 0.0;
            }
            // ================= flux differences, bottom-slope source, per-cell sources =================
            R dU[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                const R Fp = shfl_up(F[k]);                            // face (i-1 | i)
                dU[k] = fma(p.nrdy, G[k] - Gl[k], p.nrdx * (F[k] - Fp));
            }
            if (HASB) {
                dU[1] = fma(p.botx * (tBr - tBl), Lx[0] + Rx[0], dU[1]);
                dU[2] = fma(p.boty * (ty1 - tyj), hsy, dU[2]);
            }
            dU[0] += src;
            if (has_factor) {
                const double fc = *(FAST || j < Ny ? pfac : a.factor);
                pfac += Nx;
                dU[0] = fma(fc, -infil_t, dU[0]);
            }
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                dprev[k] = fwd[k];
                Ry0[k] = Ry1[k]; Gl[k] = G[k];
            }
            hsy = Ly1[0] + Ry1[0];
            // ================= store =================
            const bool st = out_lane && (FAST || j < jb);
            if (STAGE == 1) {
                if (st) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) *reinterpret_cast<R*>(ps + k * planeB) = dU[k];
                }
            } else {
                // rungekutta.jl:39-49 stage 2, then friction (prev = U^1) and the dry clean-up
                R o[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const R un = *reinterpret_cast<const R*>(smU + uU[0] + Rr * U_SLOT + k * U_VAR);
                    o[k] = 0.5 * (un + fma(dt, dU[k], q0[k]));
                }
                const R fac = tail_factor1<FRIC>(p, q0, bcell(tyj, ty1), czdt, o[0]);
                o[1] *= fac;
                o[2] *= fac;
                if (st) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) *reinterpret_cast<R*>(ps + k * planeB) = o[k];
                }
            }
            ps += rowB;
            // corner state of the next row
            Ej = E1;
            tyj = ty1;
            ty2 = ty3;
        }
    };
    for (int j0 = ja; j0 < jb; j0 += 3, bank ^= 1) {
        if (j0 + 6 < Ny - 2) trip(std::true_type{}, j0);
        else trip(std::false_type{}, j0);
    }
    cp_async_wait_all();                                  // the two rows fetched past the chunk: drain before the rings are reused
  }   // work queue

    if (STAGE == 1) {
        // CFL reduction: warp shuffle -> one atomicMax per warp and direction (non-negative bit patterns)
        unsigned long long bx = Bits<R>::enc(msx), by = Bits<R>::enc(msy);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long tx = __shfl_xor_sync(0xffffffffu, bx, o);
            const unsigned long long ty = __shfl_xor_sync(0xffffffffu, by, o);
            bx = tx > bx ? tx : bx;
            by = ty > by ? ty : by;
        }
        if (lane == 0) {
            if (bx) atomicMax(&a.ctl->maxspeed_bits[0], bx);
            if (by) atomicMax(&a.ctl->maxspeed_bits[1], by);
        }
    }
}

}  // namespace d1
SWE_NS_END  // namespace swe

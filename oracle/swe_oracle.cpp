// =====================================================================================
// oracle/swe_oracle.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// A literal, unfused CPU restatement of the shallow-water finite-volume step of
// sintefmath/VolumeFluxes.jl (module SinFVM).  Every function cites the reference
// file:line it follows (paths relative to /root/reference/src).  The reference is pure
// Julia and no Julia toolchain exists in this container, so the reference itself cannot
// be executed here: PARITY IS PINNED BY THE REFERENCE'S OWN KNOWN-ANSWER / INVARIANT
// TESTS (ported to tests/test_oracle_*.py), not by reference-produced vectors.
//
// Precision model: templated on the storage type S (the reference's backend.realtype).
// All scalars that are Float64 in the reference (dx, theta, rho, g, cutoff, kappa, Cz,
// rain rate ...) are `double` here, and the C++ usual arithmetic conversions
// (float o float -> float, float o double -> double) coincide with Julia's promotion
// rules, so every expression below is written with the same operand types and the same
// association as the Julia source.  Values are rounded to S exactly where the reference
// stores into a realtype array.  Build with -O2 -ffp-contract=off (no FMA contraction).
// =====================================================================================
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <type_traits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

extern "C" {
struct orc_config {
    int32_t dim;          // 1 or 2
    int32_t dtype;        // 0 = Float32, 1 = Float64 (backend.realtype)
    int32_t nx, ny, gc;
    int32_t bc;           // 0 periodic, 1 wall, 2 neumann (bc.jl semantics, sequential CPU order)
    int32_t eq_kind;      // 0 = ShallowWaterEquations(1D) "practical", 1 = ...Pure
    int32_t recon;        // 0 NoReconstruction, 1 LinearReconstruction
    int32_t stepper;      // 0 ForwardEulerStepper, 1 RungeKutta2
    int32_t friction;     // 0 none, 1 bh2021, 2 bsa2012, 3 fcg2016
    int32_t has_bathymetry;   // 0: ConstantBottomTopography(const_B), 1: BottomTopography array
    int32_t has_bottom_source;
    int32_t rain_kind;    // 0 none, 1 ConstantRain, 2 TimeDependentRain, 3 FunctionalRain (per-cell values of the closure)
    int32_t infil_kind;   // 0 none, 1 ConstantInfiltration, 2 HortonInfiltration
    int32_t infil_first;  // order of the two per-cell source terms in the source list
    int32_t n_rain;       // TimeDependentRain table length
    double extent[4];
    double theta, rho, g, depth_cutoff, kappa, cfl, Cz, const_B, t0;
    double rain_rate, infil_const, f0, fc, k;
    double rain_rates[16], rain_times[16];
};
}

namespace {

// ---- Julia Base.max / Base.min on floats: NaN-propagating, signed-zero aware -----------
inline double jmax(double x, double y) {
    if (std::isnan(x) || std::isnan(y)) return std::numeric_limits<double>::quiet_NaN();
    if (x == y) return std::signbit(x) ? y : x;   // max(-0.0, 0.0) == 0.0
    return x > y ? x : y;
}
inline double jmin(double x, double y) {
    if (std::isnan(x) || std::isnan(y)) return std::numeric_limits<double>::quiet_NaN();
    if (x == y) return std::signbit(x) ? x : y;   // min(-0.0, 0.0) == -0.0
    return x < y ? x : y;
}
inline double jmax3(double a, double b, double c) { return jmax(jmax(a, b), c); }
inline double jmin3(double a, double b, double c) { return jmin(jmin(a, b), c); }

// reconstruction/linearreconstruction.jl:28-35
inline double minmod(double a, double b, double c) {
    if (a > 0 && b > 0 && c > 0) return jmin3(a, b, c);
    if (a < 0 && b < 0 && c < 0) return jmax3(a, b, c);
    return 0.0;
}

// reconstruction/linearreconstruction.jl:37-43 (per component; differences stay in S)
template <class S>
inline double minmod_slope(S left, S center, S right, double theta) {
    S forward_diff = right - center;
    S backward_diff = center - left;
    double central_diff = (forward_diff + backward_diff) / 2.0;
    return minmod(theta * forward_diff, central_diff, theta * backward_diff);
}

struct Vec3 { double v[3]; };

template <class S>
struct Oracle {
    orc_config c;
    int nvar;                 // 2 (1D) or 3 (2D)
    int64_t Nx, Ny, plane;    // total cells; plane = Nx*Ny
    double dx, dy;
    S cfl;                    // simulator.jl:66 cfl::FloatType
    S t, dt;                  // simulator.jl:65,67 MVector{1,FloatType}
    std::vector<S> vol[3];    // substep_outputs (simulator.jl:86)
    std::vector<S> left, right, ws;   // system.jl:40-42
    std::vector<S> B;         // corner / interface bathymetry, (Nx+1)[x(Ny+1)]
    std::vector<double> factor;  // Horton factor (Nx,Ny); Float64 in the reference default ctor
    std::vector<double> rainp;   // FunctionalRain: rain_function(t, x, y) at every cell centre (source_term_rain.jl:65-68), host-evaluated
    int cur[3];               // rotation of substep_outputs (simulator.jl:175-176)

    explicit Oracle(const orc_config& cfg) : c(cfg) {
        nvar = c.dim == 1 ? 2 : 3;
        Nx = c.nx + 2 * c.gc;                       // grid.jl:75-76, 86-87
        Ny = c.dim == 1 ? 1 : c.ny + 2 * c.gc;
        plane = Nx * Ny;
        dx = (c.extent[1] - c.extent[0]) / c.nx;    // grid.jl:73-74, 83-85
        dy = c.dim == 1 ? 0.0 : (c.extent[3] - c.extent[2]) / c.ny;
        cfl = (S)c.cfl;
        t = (S)c.t0;
        dt = (S)0;
        for (int i = 0; i < 3; ++i) { vol[i].assign(plane * nvar, (S)0); cur[i] = i; }
        left.assign(plane * nvar, (S)0);
        right.assign(plane * nvar, (S)0);
        ws.assign(plane, (S)0);
        if (c.has_bathymetry) B.assign((Nx + 1) * (c.dim == 1 ? 1 : Ny + 1), (S)0);
        if (c.infil_kind == 2) factor.assign(plane, 1.0);
        if (c.rain_kind == 3) rainp.assign(plane, 0.0);
    }
    int nsub() const { return c.stepper == 1 ? 2 : 1; }   // rungekutta.jl:24, forwardeuler.jl:26

    // ---- layout: volume.jl / buffer.jl:29-42 : data[i,j,v], column-major, 0-based here --
    inline int64_t at(int64_t i, int64_t j, int v) const { return i + Nx * j + plane * v; }
    inline int64_t atB(int64_t i, int64_t j) const { return i + (Nx + 1) * j; }

    // ---- bottom_topography.jl:87-101 (0-based i,j) -----------------------------------
    inline double B_cell(int64_t i, int64_t j) const {
        if (!c.has_bathymetry) return c.const_B;
        if (c.dim == 1) return 0.5 * (B[i] + B[i + 1]);                       // :92
        return 0.25 * (B[atB(i, j)] + B[atB(i + 1, j)] + B[atB(i, j + 1)] + B[atB(i + 1, j + 1)]);  // :97
    }
    inline double B_face_left(int64_t i, int64_t j, int dir) const {
        if (!c.has_bathymetry) return c.const_B;
        if (c.dim == 1) return B[i];                                          // :93
        if (dir == 0) return 0.5 * (B[atB(i, j)] + B[atB(i, j + 1)]);         // :98
        return 0.5 * (B[atB(i, j)] + B[atB(i + 1, j)]);                       // :100
    }
    inline double B_face_right(int64_t i, int64_t j, int dir) const {
        if (!c.has_bathymetry) return c.const_B;
        if (c.dim == 1) return B[i + 1];                                      // :94
        if (dir == 0) return 0.5 * (B[atB(i + 1, j)] + B[atB(i + 1, j + 1)]); // :99
        return 0.5 * (B[atB(i, j + 1)] + B[atB(i + 1, j + 1)]);               // :101
    }
    inline double eqB_cell(int64_t i, int64_t j) const {                      // equation.jl:86-87
        return c.eq_kind == 1 ? 0.0 : B_cell(i, j);
    }

    // ---- equations/equation.jl:38-62 --------------------------------------------------
    template <class H>
    inline double desingularize(H h) const {
        if (c.eq_kind == 1) return h;                                         // :38
        const double kappa = c.kappa;
        H h2 = h * h;                                                         // h^2 in typeof(h)
        double inner = jmin(h2 / (2 * kappa) + kappa / 2.0, kappa);
        double mag = jmax((double)std::fabs(h), inner);
        H sgn = std::copysign((H)1, h);                                       // copysign(1, h)
        return sgn * mag;                                                     // :56
    }

    // ---- physical flux + eigenvalues ---------------------------------------------------
    // swe_2D.jl:43-81, swe_2D_pure.jl:29-63, swe_1D.jl:44-58, swe_1D_pure.jl:28-41
    inline void flux_eig(int dir, S h, S hu, S hv, Vec3& F, double& e1, double& e2) const {
        const double rho = c.rho, g = c.g;
        if (c.eq_kind == 1) {                       // Pure: no desingularisation, hu/h in S
            S hn = dir == 0 ? hu : hv;              // normal momentum
            if (c.dim == 1) {
                F.v[0] = rho * hu;
                F.v[1] = rho * hu * hu / h + 0.5 * rho * g * (h * h);
                F.v[2] = 0.0;
            } else if (dir == 0) {
                F.v[0] = rho * hu;
                F.v[1] = rho * hu * hu / h + 0.5 * rho * g * (h * h);
                F.v[2] = rho * hu * hv / h;
            } else {
                F.v[0] = rho * hv;
                F.v[1] = rho * hu * hv / h;
                F.v[2] = rho * hv * hv / h + 0.5 * rho * g * (h * h);
            }
            S un = hn / h;
            e1 = un + std::sqrt(g * h);
            e2 = un - std::sqrt(g * h);
            return;
        }
        const double hstar = desingularize(h);
        const double u = hu / hstar;                // desingularize(eq, h, hu) equation.jl:64-67
        const double v = c.dim == 1 ? 0.0 : hv / hstar;
        if (c.dim == 1) {
            F.v[0] = rho * h * u;
            F.v[1] = rho * h * u * u + 0.5 * rho * g * (h * h);
            F.v[2] = 0.0;
        } else if (dir == 0) {
            F.v[0] = rho * h * u;
            F.v[1] = rho * h * u * u + 0.5 * rho * g * (h * h);
            F.v[2] = rho * h * u * v;
        } else {
            F.v[0] = rho * h * v;
            F.v[1] = rho * h * u * v;
            F.v[2] = rho * h * v * v + 0.5 * rho * g * (h * h);
        }
        const double un = dir == 0 ? u : v;
        e1 = un + std::sqrt(g * h);
        e2 = un - std::sqrt(g * h);
    }

    // ---- numericalflux/swe/centralupwind.jl:32-45 (generic) and :48-77 (practical) -----
    inline void central_upwind(int dir, const S* fm, const S* fp, Vec3& F, double& speed) const {
        Vec3 Fm{{0, 0, 0}}, Fp{{0, 0, 0}};
        double em1 = 0, em2 = 0, ep1 = 0, ep2 = 0;
        const bool practical = c.eq_kind == 0;
        if (!practical || fm[0] > c.depth_cutoff) flux_eig(dir, fm[0], fm[1], fm[2], Fm, em1, em2);
        if (!practical || fp[0] > c.depth_cutoff) flux_eig(dir, fp[0], fp[1], fp[2], Fp, ep1, ep2);
        const double aplus = jmax3(ep1, em1, 0.0);
        const double aminus = jmin3(ep2, em2, 0.0);
        if (practical && std::fabs(aplus - aminus) < c.kappa) {              // :67-69
            F = Vec3{{0, 0, 0}};
            speed = 0.0;
            return;
        }
        for (int k = 0; k < nvar; ++k) {
            S jump = fp[k] - fm[k];                                           // faceplus .- faceminus in S
            F.v[k] = (aplus * Fm.v[k] - aminus * Fp.v[k]) / (aplus - aminus) +
                     ((aplus * aminus) / (aplus - aminus)) * jump;            // :43 / :71
        }
        if (nvar == 2) F.v[2] = 0.0;
        if (practical && fm[0] < c.depth_cutoff && fp[0] < c.depth_cutoff) speed = 0.0;   // :73-75
        else speed = jmax(std::fabs(aplus), std::fabs(aminus));
    }

    inline void load(const std::vector<S>& a, int64_t i, int64_t j, S* q) const {
        for (int k = 0; k < nvar; ++k) q[k] = a[at(i, j, k)];
        if (nvar == 2) q[2] = (S)0;
    }

    // ---- iteration spaces: grid.jl:213-253, kernel_abstractions.jl:66-85, 123-131 -------
    // for_each_inner_cell(dir; ghostcells=g): active index in [g, N_d-g) (0-based),
    // passive index over the interior only.
    void inner_range(int dir, int g, int64_t& i0, int64_t& i1, int64_t& j0, int64_t& j1) const {
        if (c.dim == 1) { i0 = g; i1 = Nx - g; j0 = 0; j1 = 1; return; }
        if (dir == 0) { i0 = g; i1 = Nx - g; j0 = c.gc; j1 = Ny - c.gc; }
        else          { i0 = c.gc; i1 = Nx - c.gc; j0 = g; j1 = Ny - g; }
    }

    // ---- reconstruction ---------------------------------------------------------------
    void reconstruct(const std::vector<S>& in, int dir) {
        if (c.recon == 0) {
            // noreconstruction.jl:23-46 : ALL cells
#pragma omp parallel for
            for (int64_t j = 0; j < Ny; ++j)
                for (int64_t i = 0; i < Nx; ++i) {
                    for (int k = 0; k < nvar; ++k) {
                        left[at(i, j, k)] = in[at(i, j, k)];
                        right[at(i, j, k)] = in[at(i, j, k)];
                    }
                }
            if (c.eq_kind == 0) {
#pragma omp parallel for
                for (int64_t j = 0; j < Ny; ++j)
                    for (int64_t i = 0; i < Nx; ++i) {
                        left[at(i, j, 0)] = (S)(in[at(i, j, 0)] - B_cell(i, j));
                        right[at(i, j, 0)] = (S)(in[at(i, j, 0)] - B_cell(i, j));
                    }
            }
            return;
        }
        // linearreconstruction.jl:44-56 (generic) / :58-102 (practical SWE), ghostcells=1
        int64_t i0, i1, j0, j1;
        inner_range(dir, 1, i0, i1, j0, j1);
        const int64_t di = dir == 0 ? 1 : 0, dj = dir == 0 ? 0 : 1;
#pragma omp parallel for
        for (int64_t j = j0; j < j1; ++j)
            for (int64_t i = i0; i < i1; ++i) {
                S ql[3], qc[3], qr[3];
                load(in, i - di, j - dj, ql);
                load(in, i, j, qc);
                load(in, i + di, j + dj, qr);
                double slope[3] = {0, 0, 0};
                for (int k = 0; k < nvar; ++k) slope[k] = minmod_slope(ql[k], qc[k], qr[k], c.theta);
                if (c.eq_kind == 0) {
                    const double Bl = B_face_left(i, j, dir);
                    const double Br = B_face_right(i, j, dir);
                    const S w = qc[0];
                    if (w - 0.5 * slope[0] < Bl) slope[0] = 2.0 * (w - Bl);          // :81-84
                    else if (w + 0.5 * slope[0] < Br) slope[0] = 2.0 * (Br - w);     // :86-89
                    for (int k = 0; k < nvar; ++k) {
                        left[at(i, j, k)] = (S)(qc[k] - 0.5 * slope[k]);             // :94
                        right[at(i, j, k)] = (S)(qc[k] + 0.5 * slope[k]);            // :95
                    }
                    left[at(i, j, 0)] = (S)(left[at(i, j, 0)] - Bl);                 // :98
                    right[at(i, j, 0)] = (S)(right[at(i, j, 0)] - Br);               // :99
                } else {
                    for (int k = 0; k < nvar; ++k) {
                        left[at(i, j, k)] = (S)(qc[k] - 0.5 * slope[k]);             // :50
                        right[at(i, j, k)] = (S)(qc[k] + 0.5 * slope[k]);            // :51
                    }
                }
            }
    }

    // ---- numericalflux/numericalflux.jl:21-33 -------------------------------------------
    S compute_flux(std::vector<S>& out, int dir) {
        const double delta = dir == 0 ? dx : dy;
        int64_t i0, i1, j0, j1;
        inner_range(dir, c.gc, i0, i1, j0, j1);
        const int64_t di = dir == 0 ? 1 : 0, dj = dir == 0 ? 0 : 1;
#pragma omp parallel for
        for (int64_t j = j0; j < j1; ++j)
            for (int64_t i = i0; i < i1; ++i) {
                S Rm[3], Lr[3], Rl[3], Lm[3];
                load(right, i, j, Rm);
                load(left, i + di, j + dj, Lr);
                load(right, i - di, j - dj, Rl);
                load(left, i, j, Lm);
                Vec3 Fr, Fl;
                double sr, sl;
                central_upwind(dir, Rm, Lr, Fr, sr);        // F(right[imiddle], left[iright])
                central_upwind(dir, Rl, Lm, Fl, sl);        // F(right[ileft], left[imiddle])
                for (int k = 0; k < nvar; ++k)
                    out[at(i, j, k)] = (S)(out[at(i, j, k)] - 1 / delta * (Fr.v[k] - Fl.v[k]));   // :27
                ws[at(i, j, 0)] = (S)jmax(sr, sl);                                             // :28
            }
        S m = ws[0];
#pragma omp parallel for reduction(max : m)
        for (int64_t q = 1; q < plane; ++q) m = (S)jmax(m, ws[q]);   // maximum(wavespeeds) :32
        return m;
    }

    // ---- sourceterms/source_term_bottom.jl:24-41 ----------------------------------------
    void bottom_source(std::vector<S>& out, int dir) {
        const double delta = dir == 0 ? dx : dy;
        const int mom = dir == 0 ? 1 : 2;
        int64_t i0, i1, j0, j1;
        inner_range(dir, c.gc, i0, i1, j0, j1);
#pragma omp parallel for
        for (int64_t j = j0; j < j1; ++j)
            for (int64_t i = i0; i < i1; ++i) {
                const double Br = B_face_right(i, j, dir);
                const double Bl = B_face_left(i, j, dir);
                S hsum = right[at(i, j, 0)] + left[at(i, j, 0)];
                out[at(i, j, mom)] = (S)(out[at(i, j, mom)] + -c.g * ((Br - Bl) / delta) * (hsum / 2.0));  // :38
            }
    }

    // ---- source_term_rain.jl:30,45-54,71-76 ; source_term_infiltration.jl:68-88 ----------
    double rain_rate(double tt) const {
        if (c.rain_kind == 1) return c.rain_rate / 3600;
        int index = 0;
        while (index < c.n_rain - 1) {
            if (tt < c.rain_times[index + 1]) break;
            index += 1;
        }
        return c.rain_rates[index] / 3600;
    }
    void point_sources(std::vector<S>& out, S tt) {
        for (int pass = 0; pass < 2; ++pass) {
            const bool do_infil = (pass == 0) == (c.infil_first != 0);
            if (do_infil && c.infil_kind != 0) {
#pragma omp parallel for
                for (int64_t q = 0; q < plane; ++q) {
                    double f = c.infil_kind == 1 ? c.infil_const
                                                 : factor[q] * (c.fc + (c.f0 - c.fc) * std::exp(-c.k * tt));
                    out[q] = (S)(out[q] - f);
                }
            }
            if (!do_infil && c.rain_kind == 3) {        // output_h[index] += rain_function(t, x, y)  (:65-68, 73-75)
#pragma omp parallel for
                for (int64_t q = 0; q < plane; ++q) out[q] = (S)(out[q] + rainp[q]);
            } else if (!do_infil && c.rain_kind != 0) {
                const double r = rain_rate(tt);
#pragma omp parallel for
                for (int64_t q = 0; q < plane; ++q) out[q] = (S)(out[q] + r);
            }
        }
    }

    // ---- system.jl:62-75 ----------------------------------------------------------------
    void add_time_derivative(std::vector<S>& out, const std::vector<S>& state, S tt, S maxspeed[2]) {
        maxspeed[0] = maxspeed[1] = (S)0;
        for (int dir = 0; dir < c.dim; ++dir) {
            reconstruct(state, dir);
            maxspeed[dir] = compute_flux(out, dir);
            if (c.has_bottom_source) bottom_source(out, dir);
        }
        point_sources(out, tt);
    }

    // ---- friction.jl:30-84 ----------------------------------------------------------------
    double friction_fn(double h, double speed) const {
        double denom;
        if (c.friction == 1) denom = h * h;                    // bh2021 :40-43
        else if (c.friction == 2) denom = std::cbrt(h) * h;    // bsa2012 :30-33
        else denom = std::cbrt(h) * h * h;                     // fcg2016 :35-38
        return -c.Cz * speed / denom;
    }
    void implicit_friction(std::vector<S>& out, const std::vector<S>& prev, S dts) {
        if (c.friction == 0) return;
#pragma omp parallel for
        for (int64_t j = 0; j < Ny; ++j)
            for (int64_t i = 0; i < Nx; ++i) {
                const double Bm = eqB_cell(i, j);
                const double hstar = desingularize(prev[at(i, j, 0)] - Bm);
                const double u = prev[at(i, j, 1)] / hstar;
                double speed;
                if (c.dim == 1) speed = std::sqrt(u * u);
                else {
                    const double v = prev[at(i, j, 2)] / hstar;
                    speed = std::sqrt(u * u + v * v);
                }
                if (speed == 0.0) speed = 0.0;
                const double fs = friction_fn(hstar, speed);
                out[at(i, j, 0)] = (S)(out[at(i, j, 0)] / (1 - dts * 0.0));
                for (int k = 1; k < nvar; ++k)
                    out[at(i, j, k)] = (S)(out[at(i, j, k)] / (1 - dts * fs));
            }
    }

    // ---- timestepper/timestepper.jl:25-51 -------------------------------------------------
    void post_proc(std::vector<S>& out) {
        if (c.eq_kind != 0) return;
#pragma omp parallel for
        for (int64_t j = 0; j < Ny; ++j)
            for (int64_t i = 0; i < Nx; ++i) {
                const double b = B_cell(i, j);
                if (out[at(i, j, 0)] - b < c.depth_cutoff) {
                    out[at(i, j, 0)] = (S)jmax((double)out[at(i, j, 0)], b);
                    for (int k = 1; k < nvar; ++k) out[at(i, j, k)] = (S)0.0;
                }
            }
    }

    // ---- bc.jl:22-114 ; ghost iteration space grid.jl:246-253 ------------------------------
    // lo / hi: apply the BC on the low / high side (false on a y-strip side owned by a neighbour rank)
    void update_bc_dir(std::vector<S>& d, int dir, bool lo = true, bool hi = true) {
        const int64_t gc = c.gc;
        const int64_t N = dir == 0 ? Nx : Ny;         // totalcells[dir]
        int64_t p0, p1;                                // passive range (interior)
        if (c.dim == 1) { p0 = 0; p1 = 1; }
        else if (dir == 0) { p0 = gc; p1 = Ny - gc; }
        else { p0 = gc; p1 = Nx - gc; }
        auto idx = [&](int64_t a1 /*1-based active*/, int64_t p, int k) {
            return dir == 0 ? at(a1 - 1, p, k) : at(p, a1 - 1, k);
        };
        const int normal = dir == 0 ? 1 : 2;
        for (int64_t p = p0; p < p1; ++p)
            for (int64_t g = 1; g <= gc; ++g) {        // sequential order: g fastest (CPU backend)
                for (int k = 0; k < nvar; ++k) {
                    if (c.bc == 0) {                   // PeriodicBC :22-40
                        if (lo) d[idx(g, p, k)] = d[idx(N + g - 2 * gc, p, k)];
                        if (hi) d[idx(N - (gc - g), p, k)] = d[idx(gc + g, p, k)];
                    } else if (c.bc == 2) {            // NeumannBC :42-60 (reference quirks kept)
                        if (lo) {
                            if (c.dim == 1 || dir == 0) d[idx(g, p, k)] = d[idx(2 * gc - g + 1, p, k)];
                            else d[idx(g, p, k)] = d[idx(N + g - 2 * gc, p, k)];
                        }
                        if (hi) d[idx(N - (gc - g), p, k)] = d[idx(N - g + 1, p, k)];
                    } else {                           // WallBC :63-110
                        const S sgn = (k == normal) ? (S)-1 : (S)1;
                        if (lo) d[idx(g, p, k)] = sgn * d[idx(gc * 2 + 1 - g, p, k)];
                        if (hi) d[idx(N - gc + g, p, k)] = sgn * d[idx(N - gc - g + 1, p, k)];
                    }
                }
            }
    }
    void update_bc(std::vector<S>& d) {
        for (int dir = 0; dir < c.dim; ++dir) update_bc_dir(d, dir);
    }

    // ---- timestepper/rungekutta.jl:26-52, forwardeuler.jl:28-47 ----------------------------
    // returns the (unrounded, Float64) dt exactly like the Julia local variable
    double do_substep(std::vector<S>& out, int substep /*0-based*/, double dt_in, double max_dt, S tt) {
        const int64_t n = plane * nvar;
#pragma omp parallel for
        for (int64_t q = 0; q < n; ++q) out[q] = (S)0;
        const std::vector<S>& state = vol[cur[substep]];
        S wavespeed[2];
        add_time_derivative(out, state, tt, wavespeed);
        double dtl = dt_in;
        if (substep == 0) {
            // simulator.jl:147-150 timestep_computer
            double ddt = dx / wavespeed[0];
            if (c.dim == 2) ddt = jmin(ddt, dy / wavespeed[1]);
            dtl = jmin(cfl * ddt, max_dt);
        }
        if (substep == 0) {
            // dt is the freshly computed Float64 local (rungekutta.jl:35,40)
#pragma omp parallel for
            for (int64_t q = 0; q < n; ++q) out[q] = (S)(state[q] + dtl * out[q]);
        } else {
            // dt is the argument = simulator.current_timestep[1], a realtype value: the whole
            // expression stays in realtype (simulator.jl:151,155; rungekutta.jl:40)
            const S dts = (S)dt_in;
#pragma omp parallel for
            for (int64_t q = 0; q < n; ++q) out[q] = (S)(state[q] + dts * out[q]);
        }
        if (c.stepper == 1 && substep == 1) {
            const std::vector<S>& first = vol[cur[0]];
#pragma omp parallel for
            for (int64_t q = 0; q < n; ++q) out[q] = (S)(0.5 * (first[q] + out[q]));
        }
        return dtl;
    }

    // ---- simulator.jl:145-177 ----------------------------------------------------------------
    void perform_step(double max_dt) {
        for (int s = 0; s < nsub(); ++s) {
            std::vector<S>& out = vol[cur[s + 1]];
            dt = (S)do_substep(out, s, (double)dt, max_dt, t);
            implicit_friction(out, vol[cur[s]], dt);
            post_proc(out);
            update_bc(out);
        }
        std::swap(cur[0], cur[nsub()]);
    }

    // ---- simulator.jl:179-212 ---------------------------------------------------------------
    int64_t simulate_to_time(double endtime, int match_endtime, double maximum_timestep, int64_t max_steps) {
        int64_t n = 0;
        while (t < endtime) {
            double max_dt = match_endtime ? endtime - t : std::numeric_limits<double>::infinity();
            if (maximum_timestep > 0) max_dt = jmin(max_dt, maximum_timestep);
            perform_step(max_dt);
            t += dt;
            ++n;
            if (max_steps > 0 && n >= max_steps) break;
        }
        return n;
    }

    // ---- simulator.jl:119-139 + interior_volume.jl:177-186 -------------------------------------
    void set_state_interior(const S* const* comp) {
        std::vector<S>& d = vol[cur[0]];
        const int64_t gc = c.gc;
        const int64_t ny_i = c.dim == 1 ? 1 : c.ny;
        for (int k = 0; k < nvar; ++k)
            for (int64_t j = 0; j < ny_i; ++j)
                for (int64_t i = 0; i < c.nx; ++i)
                    d[at(i + gc, c.dim == 1 ? 0 : j + gc, k)] = comp[k][i + (int64_t)c.nx * j];
        update_bc(d);
    }
};

struct Handle {
    int dtype;
    Oracle<float>* f;
    Oracle<double>* d;
};

template <class F>
auto dispatch(void* hp, F&& fn) {
    Handle* h = (Handle*)hp;
    if (h->dtype == 0) return fn(*h->f);
    return fn(*h->d);
}

}  // namespace

extern "C" {

void* orc_create(const orc_config* cfg) {
    Handle* h = new Handle{cfg->dtype, nullptr, nullptr};
    if (cfg->dtype == 0) h->f = new Oracle<float>(*cfg);
    else h->d = new Oracle<double>(*cfg);
    return h;
}
void orc_destroy(void* hp) {
    Handle* h = (Handle*)hp;
    delete h->f;
    delete h->d;
    delete h;
}
// B has the storage dtype; size (Nx+1) [x (Ny+1)]
int orc_set_bathymetry(void* hp, const void* Bp) {
    return dispatch(hp, [&](auto& o) {
        using S = typename std::remove_reference<decltype(o.B[0])>::type;
        if (!o.c.has_bathymetry) return -1;
        std::memcpy(o.B.data(), Bp, o.B.size() * sizeof(S));
        return 0;
    });
}
int orc_set_infiltration_factor(void* hp, const double* f) {
    return dispatch(hp, [&](auto& o) {
        if (o.c.infil_kind != 2) return -1;
        std::copy(f, f + o.plane, o.factor.begin());
        return 0;
    });
}
int orc_set_rain_plane(void* hp, const double* r) {
    return dispatch(hp, [&](auto& o) {
        if (o.c.rain_kind != 3) return -1;
        std::copy(r, r + o.plane, o.rainp.begin());
        return 0;
    });
}
int orc_set_state_interior(void* hp, const void* w, const void* hu, const void* hv) {
    return dispatch(hp, [&](auto& o) {
        using S = typename std::remove_reference<decltype(o.ws[0])>::type;
        const S* comp[3] = {(const S*)w, (const S*)hu, (const S*)hv};
        o.set_state_interior(comp);
        return 0;
    });
}
int orc_set_state_full(void* hp, int which, const void* data) {
    return dispatch(hp, [&](auto& o) {
        using S = typename std::remove_reference<decltype(o.ws[0])>::type;
        std::memcpy(o.vol[o.cur[which]].data(), data, o.vol[0].size() * sizeof(S));
        return 0;
    });
}
int orc_get_state_full(void* hp, int which, void* data) {
    return dispatch(hp, [&](auto& o) {
        using S = typename std::remove_reference<decltype(o.ws[0])>::type;
        std::memcpy(data, o.vol[o.cur[which]].data(), o.vol[0].size() * sizeof(S));
        return 0;
    });
}
int orc_get_faces(void* hp, void* l, void* r) {
    return dispatch(hp, [&](auto& o) {
        using S = typename std::remove_reference<decltype(o.ws[0])>::type;
        std::memcpy(l, o.left.data(), o.left.size() * sizeof(S));
        std::memcpy(r, o.right.data(), o.right.size() * sizeof(S));
        return 0;
    });
}
int orc_get_wavespeeds(void* hp, void* w) {
    return dispatch(hp, [&](auto& o) {
        using S = typename std::remove_reference<decltype(o.ws[0])>::type;
        std::memcpy(w, o.ws.data(), o.ws.size() * sizeof(S));
        return 0;
    });
}
int orc_get_time(void* hp, double* t, double* dt) {
    return dispatch(hp, [&](auto& o) { *t = (double)o.t; *dt = (double)o.dt; return 0; });
}
int orc_set_time(void* hp, double t, double dt) {
    return dispatch(hp, [&](auto& o) {
        using S = typename std::remove_reference<decltype(o.ws[0])>::type;
        o.t = (S)t; o.dt = (S)dt; return 0;
    });
}
int orc_step(void* hp, double max_dt) {
    return dispatch(hp, [&](auto& o) { o.perform_step(max_dt); return 0; });
}
// perform_step! followed by t += dt (one iteration of the simulate_to_time loop body)
int orc_step_advance(void* hp, double max_dt) {
    return dispatch(hp, [&](auto& o) { o.perform_step(max_dt); o.t += o.dt; return 0; });
}
int64_t orc_simulate_to_time(void* hp, double endtime, int match_endtime, double maximum_timestep, int64_t max_steps) {
    return dispatch(hp, [&](auto& o) { return o.simulate_to_time(endtime, match_endtime, maximum_timestep, max_steps); });
}
int orc_update_bc(void* hp, int which) {
    return dispatch(hp, [&](auto& o) { o.update_bc(o.vol[o.cur[which]]); return 0; });
}
int orc_reconstruct(void* hp, int dir, int which) {
    return dispatch(hp, [&](auto& o) { o.reconstruct(o.vol[o.cur[which]], dir); return 0; });
}
// out_buffer is NOT zeroed (the caller decides, like the reference's compute_flux!/add_time_derivative!)
int orc_add_time_derivative(void* hp, int in_buffer, int out_buffer, double t, double* maxspeed) {
    return dispatch(hp, [&](auto& o) {
        using S = typename std::remove_reference<decltype(o.ws[0])>::type;
        S ms[2];
        o.add_time_derivative(o.vol[o.cur[out_buffer]], o.vol[o.cur[in_buffer]], (S)t, ms);
        maxspeed[0] = ms[0]; maxspeed[1] = ms[1];
        return 0;
    });
}
int orc_implicit_friction(void* hp, int out_buffer, int prev_buffer, double dt) {
    return dispatch(hp, [&](auto& o) {
        using S = typename std::remove_reference<decltype(o.ws[0])>::type;
        o.implicit_friction(o.vol[o.cur[out_buffer]], o.vol[o.cur[prev_buffer]], (S)dt);
        return 0;
    });
}
// ---- pieces of do_substep!/perform_step! exposed one by one, so that a host-side driver can run a
// y-strip decomposition (global dt, halo rows from neighbours) with exactly the reference arithmetic ----
int orc_zero(void* hp, int which) {
    return dispatch(hp, [&](auto& o) {
        using S = typename std::remove_reference<decltype(o.ws[0])>::type;
        std::fill(o.vol[o.cur[which]].begin(), o.vol[o.cur[which]].end(), (S)0);
        return 0;
    });
}
// out = state + dt*out ; rounded != 0: dt is the realtype value of stage 2 (rungekutta.jl:39-41)
int orc_axpy(void* hp, int out_buffer, int state_buffer, double dt, int rounded) {
    return dispatch(hp, [&](auto& o) {
        using S = typename std::remove_reference<decltype(o.ws[0])>::type;
        auto& out = o.vol[o.cur[out_buffer]];
        const auto& st = o.vol[o.cur[state_buffer]];
        const int64_t n = (int64_t)out.size();
        if (rounded) {
            const S dts = (S)dt;
            for (int64_t q = 0; q < n; ++q) out[q] = (S)(st[q] + dts * out[q]);
        } else {
            for (int64_t q = 0; q < n; ++q) out[q] = (S)(st[q] + dt * out[q]);
        }
        return 0;
    });
}
// out = 0.5*(first + out) (rungekutta.jl:46-48)
int orc_rk_average(void* hp, int out_buffer, int first_buffer) {
    return dispatch(hp, [&](auto& o) {
        using S = typename std::remove_reference<decltype(o.ws[0])>::type;
        auto& out = o.vol[o.cur[out_buffer]];
        const auto& f = o.vol[o.cur[first_buffer]];
        for (size_t q = 0; q < out.size(); ++q) out[q] = (S)(0.5 * (f[q] + out[q]));
        return 0;
    });
}
int orc_update_bc_sides(void* hp, int which, int xlo, int xhi, int ylo, int yhi) {
    return dispatch(hp, [&](auto& o) {
        o.update_bc_dir(o.vol[o.cur[which]], 0, xlo != 0, xhi != 0);
        if (o.c.dim == 2) o.update_bc_dir(o.vol[o.cur[which]], 1, ylo != 0, yhi != 0);
        return 0;
    });
}
// simulator.jl:175-176
int orc_rotate(void* hp) {
    return dispatch(hp, [&](auto& o) { std::swap(o.cur[0], o.cur[o.nsub()]); return 0; });
}
int orc_post_proc(void* hp, int which) {
    return dispatch(hp, [&](auto& o) { o.post_proc(o.vol[o.cur[which]]); return 0; });
}

// number of host threads for the per-cell loops (each iteration writes its own cell: results are
// independent of the thread count).  Returns the count in effect.
int orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

// ---- scalar helpers pinned by the reference's unit tests ----------------------------------
double orc_minmod(double a, double b, double c) { return minmod(a, b, c); }
double orc_B_cell(void* hp, int64_t i, int64_t j) { return dispatch(hp, [&](auto& o) { return o.B_cell(i, j); }); }
double orc_B_face_left(void* hp, int64_t i, int64_t j, int dir) { return dispatch(hp, [&](auto& o) { return o.B_face_left(i, j, dir); }); }
double orc_B_face_right(void* hp, int64_t i, int64_t j, int dir) { return dispatch(hp, [&](auto& o) { return o.B_face_right(i, j, dir); }); }
double orc_desingularize(void* hp, double h) { return dispatch(hp, [&](auto& o) { return o.desingularize(h); }); }
double orc_rain_rate(void* hp, double t) { return dispatch(hp, [&](auto& o) { return o.rain_rate(t); }); }
double orc_infiltration(void* hp, double t, int64_t i, int64_t j) {
    return dispatch(hp, [&](auto& o) {
        if (o.c.infil_kind == 1) return o.c.infil_const;
        return o.factor[i + o.Nx * j] * (o.c.fc + (o.c.f0 - o.c.fc) * std::exp(-o.c.k * t));
    });
}

// ---- index-space enumeration (1-based triples like the Julia CartesianIndex) ----------------
// for_each_inner_cell: writes (left_i,left_j, mid_i,mid_j, right_i,right_j) per visited cell
int64_t orc_enumerate_inner_cells(void* hp, int dir, int ghostcells, int64_t* out) {
    return dispatch(hp, [&](auto& o) {
        int64_t i0, i1, j0, j1, n = 0;
        o.inner_range(dir, ghostcells, i0, i1, j0, j1);
        const int64_t di = dir == 0 ? 1 : 0, dj = dir == 0 ? 0 : 1;
        for (int64_t j = j0; j < j1; ++j)
            for (int64_t i = i0; i < i1; ++i) {
                if (out) {
                    int64_t* p = out + 6 * n;
                    p[0] = i - di + 1; p[1] = j - dj + 1; p[2] = i + 1; p[3] = j + 1; p[4] = i + di + 1; p[5] = j + dj + 1;
                }
                ++n;
            }
        return n;
    });
}
int64_t orc_enumerate_ghost_cells(void* hp, int dir, int64_t* out) {
    return dispatch(hp, [&](auto& o) {
        int64_t n = 0;
        const int64_t gc = o.c.gc;
        if (o.c.dim == 1) {
            for (int64_t g = 1; g <= gc; ++g) { if (out) { out[2 * n] = g; out[2 * n + 1] = 1; } ++n; }
            return n;
        }
        if (dir == 0) {
            for (int64_t j = gc + 1; j <= o.Ny - gc; ++j)
                for (int64_t g = 1; g <= gc; ++g) { if (out) { out[2 * n] = g; out[2 * n + 1] = j; } ++n; }
        } else {
            for (int64_t g = 1; g <= gc; ++g)
                for (int64_t i = gc + 1; i <= o.Nx - gc; ++i) { if (out) { out[2 * n] = i; out[2 * n + 1] = g; } ++n; }
        }
        return n;
    });
}
int64_t orc_enumerate_cells(void* hp, int64_t* out) {
    return dispatch(hp, [&](auto& o) {
        int64_t n = 0;
        for (int64_t j = 1; j <= o.Ny; ++j)
            for (int64_t i = 1; i <= o.Nx; ++i) { if (out) { out[2 * n] = i; out[2 * n + 1] = j; } ++n; }
        return n;
    });
}

}  // extern "C"

// engine_impl.cuh -- EngineT<P>: owns the device memory of one Simulator (one y-strip on one GPU)
// and launches the kernels.  Host orchestration mirrors simulator.jl:145-212 and system.jl:62-75.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <limits>
#include <vector>

#include "engine.h"
#include <unistd.h>

#include <cstdlib>
#include <type_traits>

#include "kernels_exchange.cuh"
#include "kernels_fused.cuh"
#include "kernels_fused_x2.cuh"
#include "kernels_fused_d1.cuh"

SWE_NS_BEGIN

#define SWE_CUDA(call)                                                                          \
    do {                                                                                        \
        cudaError_t e_ = (call);                                                                \
        if (e_ != cudaSuccess) {                                                                \
            err = std::string(#call) + ": " + cudaGetErrorString(e_);                            \
            return SWE_B200_ECUDA;                                                              \
        }                                                                                       \
    } while (0)

// What one rank publishes about its exchange window (swe_b200_get_ipc_handle): 128 bytes.
struct IpcBlob {
    cudaIpcMemHandle_t handle;   // 64 B
    long long pid;               // same-process peers use the raw pointer instead of the IPC handle
    void* base;
    long long Ny, plane, vol_bytes, mbox_offset;
    int device, rank;
    char pad[128 - 64 - 8 - 8 - 32 - 8];
};
static_assert(sizeof(IpcBlob) == 128, "IpcBlob must be 128 bytes");

template <class P>
class EngineT : public Engine {
    using S = typename P::S;
    using C = typename P::C;

public:
    swe_b200_config cfg;
    Par<C> par;
    int Nx, Ny, gc;
    int Bp = 0;             // internal pitch of the corner bathymetry: Nx+1 rounded up to even (64-bit loads)
    bool x2_ok = false;     // Float32 fast policy: the packed two-column stage kernels (kernels_fused_x2.cuh)
    x2::Par2 par2;
    d1::Par1 par1;
    long long plane;
    cudaStream_t stream = nullptr;
    S* vol[3] = {nullptr, nullptr, nullptr};
    S *left = nullptr, *right = nullptr, *ws = nullptr;
    S* B = nullptr;
    S* wrapC = nullptr;     // [2 sides][gc + 1 corner rows][Bp]: bathymetry of the wrap-around image's source rows (y-strips)
    bool wrap_set[2] = {false, false};
    double* factor = nullptr;
    double* rainp = nullptr;     // FunctionalRain values per cell (rain_kind == SWE_B200_RAIN_PLANE)
    Ctl* ctl = nullptr;
    Ctl* h_ctl = nullptr;   // pinned mirror
    int cur[3] = {0, 1, 2};
    int64_t launches = 0;
    RainTable rt;
    InfilPar ip;
    bool fused_ok = false;
    bool d1_ok = false;     // Float64 fast policy: the ring / persistent-warp stage kernels (kernels_fused_d1.cuh)
    // ---- multi-GPU exchange window: the three volumes + the mailbox live in ONE allocation ----
    char* slab = nullptr;
    size_t vol_bytes = 0, mbox_offset = 0, slab_bytes = 0;
    Mailbox* mbox = nullptr;
    unsigned int* ticket = nullptr;
    bool peers_open = false;
    char* peer_base[MAX_RANKS] = {};
    long long peer_Ny[MAX_RANKS] = {}, peer_plane[MAX_RANKS] = {}, peer_vol_bytes[MAX_RANKS] = {};
    Mailbox* peer_mbox[MAX_RANKS] = {};
    bool peer_ipc[MAX_RANKS] = {};
    int lo_rank = -1, hi_rank = -1;   // ranks my bottom / top owned rows are pushed to (-1: nobody)
    bool recv_lo = false, recv_hi = false;   // do my bottom / top ghost rows come from a neighbour?
    unsigned long long xseq = 0;        // halo exchange stamp counter (identical on all ranks)
    unsigned long long sseq = 0;        // step sequence for the CFL all-gather
    unsigned long long pending_U = 0;   // stamp of the last U exchange the next K1 must wait for
    unsigned long long bseq = 0;        // device-side rendezvous counter (k_rank_barrier)
    // the tail kernel of a fused step already ran ctl_begin for the NEXT step with these parameters: a following step
    // with the same parameters starts with K1 directly (one launch less).  Anything else that touches the control
    // block (set_time, add_time_derivative, the unfused path) clears `primed`.
    struct BeginKey {
        double max_dt_host, t_end, max_timestep;
        int use_t_end, match;
        bool operator==(const BeginKey& o) const {
            // (bitwise-equal doubles; +inf == +inf)
            return max_dt_host == o.max_dt_host && t_end == o.t_end && max_timestep == o.max_timestep && use_t_end == o.use_t_end && match == o.match;
        }
    };
    bool primed = false;
    BeginKey primed_key{};
    std::vector<cudaEvent_t> ev;   // per-stage-kernel timing events (timed runs only)
    std::vector<int> ev_stage;
    bool timing_on = false;
    bool trace_on = false;                       // SWE_B200_TRACE=1: per-launch timeline of the timed steps (stderr)
    std::vector<cudaEvent_t> tr_ev;
    std::vector<const char*> tr_label;
    void mark(const char* label) {
        if (!(timing_on && trace_on)) return;
        cudaEvent_t e;
        cudaEventCreate(&e);
        cudaEventRecord(e, stream);
        tr_ev.push_back(e);
        tr_label.push_back(label);
    }

    int init(const swe_b200_config& c) {
        cfg = c;
        gc = c.gc;
        Nx = c.nx + 2 * gc;
        Ny = c.ny + 2 * gc;
        plane = (long long)Nx * Ny;
        Bp = (Nx + 2) / 2 * 2;
        SWE_CUDA(cudaSetDevice(c.device));
        SWE_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        const int gny = c.global_ny > 0 ? c.global_ny : c.ny;
        const double dx = (c.extent[1] - c.extent[0]) / c.nx;       // grid.jl:83-85
        const double dy = (c.extent[3] - c.extent[2]) / gny;
        par.Nx = Nx; par.Ny = Ny; par.gc = gc; par.plane = plane;
        par.dx = (C)dx; par.dy = (C)dy; par.rdx = (C)(1 / dx); par.rdy = (C)(1 / dy);
        par.theta = (C)c.theta; par.rho = (C)c.rho; par.g = (C)c.g;
        par.half_rho_g = (C)(0.5 * c.rho * c.g);
        par.cutoff = (C)c.depth_cutoff; par.kappa = (C)c.kappa;
        par.two_kappa = (C)(2 * c.kappa); par.inv_two_kappa = (C)(1.0 / (2 * c.kappa));
        par.half_kappa = (C)(c.kappa / 2.0);
        par.Cz = (C)c.Cz; par.constB = (C)c.const_B;
        par.bc = c.bc; par.eq_kind = c.eq_kind; par.friction = c.friction;
        par.hasB = c.has_bathymetry; par.bottom_source = c.has_bottom_source;
        // y boundaries of this strip are physical unless a neighbouring rank owns the rows beyond
        // (periodic-y across ranks is an exchange between rank 0 and rank P-1)
        const bool multi = c.nranks > 1;
        par.ylo_phys = (!multi || (c.rank == 0 && c.bc != SWE_B200_BC_PERIODIC && c.bc != SWE_B200_BC_NEUMANN)) ? 1 : 0;
        par.yhi_phys = (!multi || (c.rank == c.nranks - 1 && c.bc != SWE_B200_BC_PERIODIC)) ? 1 : 0;
        vol_bytes = (sizeof(S) * plane * 3 + 255) / 256 * 256;
        mbox_offset = 3 * vol_bytes;
        slab_bytes = mbox_offset + 1024;
        SWE_CUDA(cudaMalloc(&slab, slab_bytes));
        SWE_CUDA(cudaMemsetAsync(slab, 0, slab_bytes, stream));
        for (int k = 0; k < 3; ++k) vol[k] = (S*)(slab + k * vol_bytes);
        mbox = (Mailbox*)(slab + mbox_offset);
        ticket = (unsigned int*)(slab + mbox_offset + 768);
        static_assert(sizeof(Mailbox) <= 768, "mailbox too large");
        if (multi) {
            if (c.nranks > MAX_RANKS) { err = "at most 8 ranks"; return SWE_B200_EUNSUPPORTED; }
            const bool per = c.bc == SWE_B200_BC_PERIODIC, neu = c.bc == SWE_B200_BC_NEUMANN;
            // periodic: the first / last strips are neighbours too.  NeumannBC (bc.jl:56-58): the low-y ghosts of the
            // domain are the image of its TOP interior rows -- a one-directional wrap-around from the last strip to the
            // first -- and its high-y ghosts evolve on their own on the last strip (k_neumann_*): nothing comes back.
            lo_rank = c.rank > 0 ? c.rank - 1 : (per ? c.nranks - 1 : -1);
            hi_rank = c.rank < c.nranks - 1 ? c.rank + 1 : ((per || neu) ? 0 : -1);
            recv_lo = c.rank > 0 || per || neu;
            recv_hi = c.rank < c.nranks - 1 || per;
        }
        if (c.has_bathymetry) {
            SWE_CUDA(cudaMalloc(&B, sizeof(S) * (size_t)Bp * (Ny + 1)));
            SWE_CUDA(cudaMemsetAsync(B, 0, sizeof(S) * (size_t)Bp * (Ny + 1), stream));
        }
        if (c.rain_kind == SWE_B200_RAIN_PLANE) {
            SWE_CUDA(cudaMalloc(&rainp, sizeof(double) * plane));
            SWE_CUDA(cudaMemsetAsync(rainp, 0, sizeof(double) * plane, stream));
        }
        if (c.infil_kind == SWE_B200_INFIL_HORTON) {
            SWE_CUDA(cudaMalloc(&factor, sizeof(double) * plane));
            std::vector<double> ones((size_t)plane, 1.0);   // HortonInfiltration default factor=1.0
            SWE_CUDA(cudaMemcpyAsync(factor, ones.data(), sizeof(double) * plane, cudaMemcpyHostToDevice, stream));
            SWE_CUDA(cudaStreamSynchronize(stream));
        }
        SWE_CUDA(cudaMalloc(&ctl, sizeof(Ctl)));
        SWE_CUDA(cudaMallocHost(&h_ctl, sizeof(Ctl)));
        std::memset(h_ctl, 0, sizeof(Ctl));
        h_ctl->t = (double)(S)c.t0;
        SWE_CUDA(cudaMemcpyAsync(ctl, h_ctl, sizeof(Ctl), cudaMemcpyHostToDevice, stream));
        rt.kind = c.rain_kind; rt.n = c.n_rain; rt.rate = c.rain_rate;
        for (int q = 0; q < 16; ++q) { rt.rates[q] = c.rain_rates[q]; rt.times[q] = c.rain_times[q]; }
        ip.kind = c.infil_kind; ip.cst = c.infil_const; ip.f0 = c.f0; ip.fc = c.fc; ip.k = c.k;
        // the fused stage kernels cover the practical 2D SWE with gc == 2 and periodic / wall BCs
        fused_ok = c.fused && c.eq_kind == 0 && gc == 2 && c.nx >= 4 && c.ny >= 4;
        setup_x2(c);
        setup_d1(c);
        if (multi && !fused_ok) {
            err = "y-strip decomposition needs the fused path (ShallowWaterEquations, gc == 2, periodic / wall BC)";
            return SWE_B200_EUNSUPPORTED;
        }
        SWE_CUDA(cudaStreamSynchronize(stream));
        return SWE_B200_OK;
    }

    ~EngineT() override {
        for (int r = 0; r < MAX_RANKS; ++r)
            if (peer_ipc[r] && peer_base[r]) cudaIpcCloseMemHandle(peer_base[r]);
        cudaFree(slab);
        cudaFree(left); cudaFree(right); cudaFree(ws); cudaFree(B); cudaFree(wrapC); cudaFree(factor); cudaFree(rainp); cudaFree(ctl); cudaFree(d_items[0]); cudaFree(d_items[1]); cudaFree(diag_buf);
        if (h_ctl) cudaFreeHost(h_ctl);
        for (auto e : ev) cudaEventDestroy(e);
        cudaFree(snap);
        if (snap_ready) cudaEventDestroy(snap_ready);
        if (snap_done) cudaEventDestroy(snap_done);
        if (copy_stream) cudaStreamDestroy(copy_stream);
        if (stream) cudaStreamDestroy(stream);
    }

    // Float32 fast policy only: parameters of the packed two-column kernels.  Needs an even row pitch
    // (64-bit loads), cutoff >= 0 and 0 < kappa < 1 (input-side dry mask; desingularisation by a saturating add).
    void setup_x2(const swe_b200_config& c) {
        x2_ok = false;
        if (!std::is_same<P, Policy<float, float, false>>::value) return;
        const char* off = std::getenv("SWE_B200_NO_X2");
        if (off && off[0] == '1') return;
        if (!fused_ok || (Nx & 1) || !(c.depth_cutoff >= 0) || !(c.kappa > 0) || !(c.kappa < 1)) return;
        if (c.rain_kind == SWE_B200_RAIN_PLANE) return;      // per-cell rain: generic fused kernels
        const int gny = c.global_ny > 0 ? c.global_ny : c.ny;
        const double dx = (c.extent[1] - c.extent[0]) / c.nx, dy = (c.extent[3] - c.extent[2]) / gny;
        auto s2 = [](double v) { return make_float2((float)v, (float)v); };
        x2::Par2& q = par2;
        q.Nx = Nx; q.Ny = Ny; q.plane = plane; q.Bp = Bp;
        q.bc = c.bc; q.friction = c.friction; q.bottom_source = c.has_bottom_source;
        q.ylo_phys = par.ylo_phys; q.yhi_phys = par.yhi_phys;
        q.rain_kind = c.rain_kind; q.infil_kind = c.infil_kind; q.infil_first = c.infil_first; q.rows = x2::ROWS;
        q.cutoff = (float)c.depth_cutoff; q.kappa = (float)c.kappa; q.constB = (float)c.const_B;
        q.theta = s2(c.theta); q.hith = s2(0.5 / c.theta); q.half = s2(0.5); q.mhalf = s2(-0.5); q.quarter = s2(0.25);
        q.g = s2(c.g); q.K = s2(0.5 * c.rho * c.g); q.i2k = s2(1.0 / (2 * c.kappa)); q.hk = s2(c.kappa / 2.0);
        q.rho = s2(c.rho); q.nrdx = s2(-1.0 / dx); q.nrdy = s2(-1.0 / dy); q.prdy = s2(1.0 / dy);
        q.botx = s2(c.has_bottom_source ? -0.25 * c.g / dx : 0.0); q.boty = s2(c.has_bottom_source ? -0.25 * c.g / dy : 0.0); q.Cz = s2(c.Cz);
        q.rows = 126;
        if (const char* e = std::getenv("SWE_B200_X2_ROWS")) q.rows = std::max(3, atoi(e) / 3 * 3);   // multiple of 3
        // (if the item tables cannot be allocated the generic fused kernels serve the configuration)
        x2_ok = build_items(c, q.rows, 1) && build_items(c, q.rows, 2);
    }
    // Float64 fast policy: parameters of the ring / persistent-warp kernels (kernels_fused_d1.cuh).  Same conditions as
    // the packed Float32 kernels except the even pitch (8-byte loads of doubles are always aligned) and kappa < 1.
    void setup_d1(const swe_b200_config& c) {
        d1_ok = false;
        if (!std::is_same<P, Policy<double, double, false>>::value) return;
        const char* off = std::getenv("SWE_B200_NO_D1");
        if (off && off[0] == '1') return;
        if (!fused_ok || !(c.depth_cutoff >= 0) || !(c.kappa > 0)) return;
        if (c.rain_kind == SWE_B200_RAIN_PLANE) return;      // per-cell rain: generic fused kernels
        const int gny = c.global_ny > 0 ? c.global_ny : c.ny;
        const double dx = (c.extent[1] - c.extent[0]) / c.nx, dy = (c.extent[3] - c.extent[2]) / gny;
        d1::Par1& q = par1;
        q.Nx = Nx; q.Ny = Ny; q.plane = plane; q.Bp = Bp;
        q.bc = c.bc; q.friction = c.friction; q.bottom_source = c.has_bottom_source;
        q.ylo_phys = par.ylo_phys; q.yhi_phys = par.yhi_phys;
        q.rain_kind = c.rain_kind; q.infil_kind = c.infil_kind; q.infil_first = c.infil_first;
        q.cutoff = c.depth_cutoff; q.kappa = c.kappa; q.constB = c.const_B;
        q.theta = c.theta; q.g = c.g; q.K = 0.5 * c.rho * c.g; q.i2k = 1.0 / (2 * c.kappa); q.rho = c.rho;
        q.nrdx = -1.0 / dx; q.nrdy = -1.0 / dy;
        q.botx = c.has_bottom_source ? -0.25 * c.g / dx : 0.0; q.boty = c.has_bottom_source ? -0.25 * c.g / dy : 0.0;
        q.Cz = c.Cz;
        int rows = 126;
        if (const char* e = std::getenv("SWE_B200_X2_ROWS")) rows = std::max(3, atoi(e) / 3 * 3);   // multiple of 3
        d1_ok = build_items(c, rows, 1, d1::VW, d1::minb<1>() * d1::WPB) && build_items(c, rows, 2, d1::VW, d1::minb<2>() * d1::WPB);
    }
    template <int STAGE>
    void launch_stage_d1(const S* U, const S* D, S* out) {
        if constexpr (std::is_same<P, Policy<double, double, false>>::value) {
            d1::Args1 a;
            a.p = par1; a.U = U; a.D = D; a.out = out; a.B = B; a.wrapLo = wrap_ptr(0); a.wrapHi = wrap_ptr(1); a.factor = factor; a.ctl = ctl; a.items = d_items[STAGE - 1];
            a.p.nitems = n_items[STAGE - 1];
            const int items = n_items[STAGE - 1];
            dim3 g(std::min((items + d1::WPB - 1) / d1::WPB, 148 * d1::minb<STAGE>()));
            dim3 b(d1::NT);
            const bool lin = cfg.recon == SWE_B200_RECON_LINEAR;
            const int fr = cfg.friction == SWE_B200_FRICTION_NONE ? 0 : (cfg.friction == SWE_B200_FRICTION_BH2021 ? 1 : 2);
#define SWE_D1_LAUNCH(R, HB, F) d1::k_stage_d1<R, HB, STAGE, (STAGE == 2 ? F : 0)><<<g, b, 0, stream>>>(a)
#define SWE_D1_FRIC(R, HB)                                                                   \
    do {                                                                                     \
        if (STAGE == 1 || fr == 0) SWE_D1_LAUNCH(R, HB, 0);                                  \
        else if (fr == 1) SWE_D1_LAUNCH(R, HB, 1);                                           \
        else SWE_D1_LAUNCH(R, HB, 2);                                                        \
    } while (0)
            if (B) {
                if (lin) SWE_D1_FRIC(1, true);
                else SWE_D1_FRIC(0, true);
            } else {
                if (lin) SWE_D1_FRIC(1, false);
                else SWE_D1_FRIC(0, false);
            }
#undef SWE_D1_FRIC
#undef SWE_D1_LAUNCH
        }
    }

    // Guided schedule of the persistent warps: every tile column (60 output columns) is cut into row segments;
    // most rows go into `big`-row items (the 4-row prologue amortises), the rest into three classes of ever
    // smaller items that are handed out LAST, sized so that each of the W resident warps of the stage gets about
    // one 63-, one 30- and one 12-row item: the launch then ends within about half a small item everywhere.
    // One table per stage: stage 2 runs 3 blocks per SM, stage 1 runs 4.
    // (SWE_B200_X2_WAVES=1 selects a variant whose big items come in whole waves of W and whose remainder is split
    // 1/2, 1/4, 1/4 over the small classes; measured 1-4 % slower on 2048..16384-row grids: more small items.)
    int4* d_items[2] = {nullptr, nullptr};
    int n_items[2] = {0, 0};
    bool build_items(const swe_b200_config& c, int big, int stage, int vw = x2::VW,
                     int resident_warps_per_sm = 0) {
        if (!resident_warps_per_sm) resident_warps_per_sm = (stage == 2 ? x2::minb<2>() : x2::minb<1>()) * x2::WPB;
        const int ntiles = (c.nx + vw - 1) / vw;
        const long long W = 148LL * resident_warps_per_sm;
        const long long L = (long long)ntiles * c.ny;
        // small grids: a big item is at most about half of a warp's whole share of tile-rows
        if (!std::getenv("SWE_B200_X2_ROWS")) big = (int)std::min<long long>(big, std::max<long long>(12, L / W / 2 / 3 * 3));
        // tiny grids (a warp's whole share is below 36 tile-rows, e.g. 1024^2: 16): the guided classes degenerate into
        // 12-row items, 1.3 per resident warp.  SWE_B200_X2_ONEWAVE=1 selects ONE wave of equal items instead, one per
        // resident warp (ceil(share) rows + one prologue).  Measured (profiles/r02j_ab_onewave.txt): no gain -- 1024^2
        // Float64 0.102 vs 0.098 ms/step, +-2 % on seven other small configurations; a 40 us launch on a grid of this
        // size is ramp, first-touch DRAM latency and the tail, not the item shape.  Off by default.
        bool one_wave = false;
        if (!std::getenv("SWE_B200_X2_ROWS") && L < 36 * W) {
            const char* e = std::getenv("SWE_B200_X2_ONEWAVE");
            if (e && e[0] == '1') {
                // rows per item: the smallest multiple of 3 whose item count (whole items per tile column) fits one wave
                for (int r = 3; r <= 48; r += 3) {
                    if ((long long)ntiles * ((c.ny + r - 1) / r) <= W) { big = r; one_wave = true; break; }
                }
            }
        }
        int sizes[4] = {big, std::min(63, std::max(12, big / 2 / 3 * 3)), std::min(30, std::max(12, big / 4 / 3 * 3)), 12};
        long long quota[4];
        double per_warp[4] = {0, 1, 1, 1};          // least items of each small class per resident warp
        if (const char* e = std::getenv("SWE_B200_X2_Q")) sscanf(e, "%lf,%lf,%lf", &per_warp[1], &per_warp[2], &per_warp[3]);
        bool waves = false;
        if (const char* e = std::getenv("SWE_B200_X2_WAVES")) waves = e[0] != '0';
        long long reserve = 0;
        for (int k = 1; k < 4; ++k) reserve += (long long)(W * per_warp[k] * sizes[k]);
        reserve = std::min(reserve, L / 2);
        if (waves) {
            const long long per_wave = W * big;
            quota[0] = (L - reserve) / per_wave * per_wave;
            const long long rest = L - quota[0];
            quota[1] = rest / 2; quota[2] = rest / 4; quota[3] = rest - quota[1] - quota[2];
        } else {
            quota[3] = std::min(L / 8, (long long)(W * per_warp[3] * sizes[3]));
            quota[2] = std::min(L / 6, (long long)(W * per_warp[2] * sizes[2]));
            quota[1] = std::min(L / 4, (long long)(W * per_warp[1] * sizes[1]));
            quota[0] = L - quota[1] - quota[2] - quota[3];
        }
        if (const char* e = std::getenv("SWE_B200_X2_FLAT")) { if (e[0] == '1') { quota[1] = quota[2] = quota[3] = 0; quota[0] = L; } }
        if (one_wave) { quota[1] = quota[2] = quota[3] = 0; quota[0] = L; }
        // rows of every tile column per class, top of the column = smallest items (any order works: the classes
        // are emitted class by class, tile-fastest, so neighbours in x run at about the same time -> L2 reuse)
        std::vector<int4> items;
        std::vector<int> pos(ntiles, 0);
        for (int k = 0; k < 4; ++k) {
            // per column share of this class, in whole items (big items rounded down: never more than whole waves); the last class takes whatever is left
            const long long per_col = (quota[k] / ntiles + (waves && k > 0 ? sizes[k] / 2 : 0)) / sizes[k] * sizes[k];
            std::vector<int> lim(ntiles);
            for (int t = 0; t < ntiles; ++t) lim[t] = k == 3 ? c.ny : (int)std::min<long long>(c.ny, pos[t] + per_col);
            bool any = true;
            while (any) {
                any = false;
                for (int t = 0; t < ntiles; ++t) {
                    if (pos[t] >= lim[t]) continue;
                    const int n = std::min(sizes[k], c.ny - pos[t]);
                    if (k < 3 && pos[t] + n > lim[t]) continue;
                    items.push_back(make_int4(t, 2 + pos[t], n, 0));
                    pos[t] += n;
                    any = true;
                }
            }
        }
        const int s = stage - 1;
        n_items[s] = (int)items.size();
        if (d_items[s]) cudaFree(d_items[s]);
        d_items[s] = nullptr;
        if (cudaMalloc(&d_items[s], sizeof(int4) * items.size()) != cudaSuccess ||
            cudaMemcpy(d_items[s], items.data(), sizeof(int4) * items.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        return true;
    }

    template <int STAGE>
    void launch_stage_x2(const S* U, const S* D, S* out) {
        if constexpr (std::is_same<P, Policy<float, float, false>>::value) {
            x2::Args2 a;
            a.p = par2; a.U = U; a.D = D; a.out = out; a.B = B; a.wrapLo = wrap_ptr(0); a.wrapHi = wrap_ptr(1); a.factor = factor; a.ctl = ctl; a.items = d_items[STAGE - 1];
            a.p.nitems = n_items[STAGE - 1];
            // persistent warps pulling (tile, rows) items: at most 4 blocks per SM are resident (128 registers)
            const int items = n_items[STAGE - 1];
            dim3 g(std::min((items + x2::WPB - 1) / x2::WPB, 148 * x2::minb<STAGE>()));
            dim3 b(x2::WPB * 32);
            const bool lin = cfg.recon == SWE_B200_RECON_LINEAR;
            const int fr = cfg.friction == SWE_B200_FRICTION_NONE ? 0 : (cfg.friction == SWE_B200_FRICTION_BH2021 ? 1 : 2);
#define SWE_X2_LAUNCH(R, HB, F, FP) x2::k_stage_x2<R, HB, STAGE, (STAGE == 2 ? F : 0), FP><<<g, b, 0, stream>>>(a)
#define SWE_X2_FACT(R, HB, F)                                                                \
    do {                                                                                     \
        if (factor) SWE_X2_LAUNCH(R, HB, F, true);                                           \
        else SWE_X2_LAUNCH(R, HB, F, false);                                                 \
    } while (0)
#define SWE_X2_FRIC(R, HB)                                                                   \
    do {                                                                                     \
        if (STAGE == 1 || fr == 0) SWE_X2_FACT(R, HB, 0);                                    \
        else if (fr == 1) SWE_X2_FACT(R, HB, 1);                                             \
        else SWE_X2_FACT(R, HB, 2);                                                          \
    } while (0)
            if (B) {
                if (lin) SWE_X2_FRIC(1, true);
                else SWE_X2_FRIC(0, true);
            } else {
                if (lin) SWE_X2_FRIC(1, false);
                else SWE_X2_FRIC(0, false);
            }
#undef SWE_X2_FRIC
#undef SWE_X2_FACT
#undef SWE_X2_LAUNCH
        }
    }

    // ------------------------------------------------------------------------------------------
    int check() {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { err = std::string("kernel launch: ") + cudaGetErrorString(e); return SWE_B200_ECUDA; }
        return SWE_B200_OK;
    }
    int ensure_face_buffers() {
        if (left) return SWE_B200_OK;
        SWE_CUDA(cudaMalloc(&left, sizeof(S) * plane * 3));
        SWE_CUDA(cudaMalloc(&right, sizeof(S) * plane * 3));
        SWE_CUDA(cudaMalloc(&ws, sizeof(S) * plane));
        SWE_CUDA(cudaMemsetAsync(left, 0, sizeof(S) * plane * 3, stream));
        SWE_CUDA(cudaMemsetAsync(right, 0, sizeof(S) * plane * 3, stream));
        SWE_CUDA(cudaMemsetAsync(ws, 0, sizeof(S) * plane, stream));
        return SWE_B200_OK;
    }
    Bathy<P> bathy() const { return Bathy<P>{B, Bp, par.constB}; }
    dim3 blk2() const { return dim3(32, 8); }
    dim3 grd2() const { return dim3((Nx + 31) / 32, (Ny + 7) / 8); }
    int grd1(long long n) const { return (int)std::min<long long>((n + 255) / 256, 148 * 16); }
    bool bad_buffer(int b) { if (b < 0 || b > 2) { err = "buffer index out of range"; return true; } return false; }

    int sync() override {
        SWE_CUDA(cudaStreamSynchronize(stream));
        return check_timeout();
    }

    int set_bathymetry(const void* Bh, int64_t n) override {
        if (!cfg.has_bathymetry) { err = "config has ConstantBottomTopography"; return SWE_B200_EINVAL; }
        if (n != (int64_t)(Nx + 1) * (Ny + 1)) {   // validate_bottom_topography, bottom_topography.jl:78-82
            err = "Bottom topography should be of size (" + std::to_string(Nx + 1) + ", " + std::to_string(Ny + 1) + ")";
            return SWE_B200_EDOMAIN;
        }
        SWE_CUDA(cudaMemcpy2DAsync(B, sizeof(S) * Bp, Bh, sizeof(S) * (Nx + 1), sizeof(S) * (Nx + 1), Ny + 1,
                                   cudaMemcpyHostToDevice, stream));
        SWE_CUDA(cudaStreamSynchronize(stream));
        return SWE_B200_OK;
    }
    // y-strips whose ghost rows are a wrap-around image (first strip under periodic / NeumannBC, last strip under periodic):
    // stage 2 rebuilds U^1 in those rows from the exchanged (U, dU) rows and needs the bathymetry of the SOURCE cells, which
    // live at the other end of the domain.  The host hands over their gc + 1 corner rows once.
    bool needs_wrap(int side) const {
        if (cfg.nranks <= 1 || !cfg.has_bathymetry) return false;
        const bool per = cfg.bc == SWE_B200_BC_PERIODIC, neu = cfg.bc == SWE_B200_BC_NEUMANN;
        return side == 0 ? (cfg.rank == 0 && (per || neu)) : (cfg.rank == cfg.nranks - 1 && per);
    }
    int set_wrap_bathymetry(int side, const void* rows, int64_t n) override {
        if (side < 0 || side > 1 || !needs_wrap(side)) { err = "this strip has no wrap-around ghost rows with bathymetry on that side"; return SWE_B200_EINVAL; }
        if (n != (int64_t)(gc + 1) * (Nx + 1)) {
            err = "Wrap bathymetry should be of size (" + std::to_string(Nx + 1) + ", " + std::to_string(gc + 1) + ")";
            return SWE_B200_EDOMAIN;
        }
        if (!wrapC) {
            SWE_CUDA(cudaMalloc(&wrapC, sizeof(S) * (size_t)Bp * (gc + 1) * 2));
            SWE_CUDA(cudaMemsetAsync(wrapC, 0, sizeof(S) * (size_t)Bp * (gc + 1) * 2, stream));
        }
        SWE_CUDA(cudaMemcpy2DAsync(wrapC + (size_t)side * (gc + 1) * Bp, sizeof(S) * Bp, rows, sizeof(S) * (Nx + 1), sizeof(S) * (Nx + 1),
                                   gc + 1, cudaMemcpyHostToDevice, stream));
        SWE_CUDA(cudaStreamSynchronize(stream));
        wrap_set[side] = true;
        return SWE_B200_OK;
    }
    const S* wrap_ptr(int side) const { return wrap_set[side] ? wrapC + (size_t)side * (gc + 1) * Bp : nullptr; }
    int set_rain_plane(const double* r, int64_t n) override {
        if (cfg.rain_kind != SWE_B200_RAIN_PLANE) { err = "no FunctionalRain (rain_kind = SWE_B200_RAIN_PLANE) configured"; return SWE_B200_EINVAL; }
        if (n != plane) {
            err = "Rain values should be of size (" + std::to_string(Nx) + ", " + std::to_string(Ny) + ")";
            return SWE_B200_EDOMAIN;
        }
        SWE_CUDA(cudaMemcpyAsync(rainp, r, sizeof(double) * n, cudaMemcpyHostToDevice, stream));
        SWE_CUDA(cudaStreamSynchronize(stream));     // the caller may reuse its buffer
        return SWE_B200_OK;
    }
    int set_infiltration_factor(const double* f, int64_t n) override {
        if (cfg.infil_kind != SWE_B200_INFIL_HORTON) { err = "no HortonInfiltration configured"; return SWE_B200_EINVAL; }
        if (n != plane) {   // validate_infiltration_factor, source_term_infiltration.jl:61-65
            err = "Infiltration factor should be of size (" + std::to_string(Nx) + ", " + std::to_string(Ny) + ")";
            return SWE_B200_EDOMAIN;
        }
        SWE_CUDA(cudaMemcpyAsync(factor, f, sizeof(double) * n, cudaMemcpyHostToDevice, stream));
        SWE_CUDA(cudaStreamSynchronize(stream));
        return SWE_B200_OK;
    }

    // interior_volume.jl:177-186 / interior_volume_variable.jl:168-176: the host arrays are dense (nx, ny) planes, the
    // volume planes are (Nx, Ny) with ghosts: strided 2D copies straight between the caller's buffers and the
    // volume (no staging allocation, no pack kernel; pinned host buffers move at PCIe speed).
    int set_state_interior(const void* w, const void* hu, const void* hv) override {
        const void* src[3] = {w, hu, hv};
        S* U = vol[cur[0]];
        for (int k = 0; k < 3; ++k)
            SWE_CUDA(cudaMemcpy2DAsync(U + (size_t)k * plane + (size_t)gc * Nx + gc, sizeof(S) * Nx, src[k], sizeof(S) * cfg.nx,
                                       sizeof(S) * cfg.nx, cfg.ny, cudaMemcpyHostToDevice, stream));
        int rc = update_bc(0);
        if (rc) return rc;
        if (peers_open) push_halo(vol[cur[0]], cur[0], 2, false);   // collective: every rank sets its state
        SWE_CUDA(cudaStreamSynchronize(stream));
        return check();
    }
    int get_state_interior(void* w, void* hu, void* hv) override {
        void* dst[3] = {w, hu, hv};
        const S* U = vol[cur[0]];
        for (int k = 0; k < 3; ++k)
            SWE_CUDA(cudaMemcpy2DAsync(dst[k], sizeof(S) * cfg.nx, U + (size_t)k * plane + (size_t)gc * Nx + gc, sizeof(S) * Nx,
                                       sizeof(S) * cfg.nx, cfg.ny, cudaMemcpyDeviceToHost, stream));
        SWE_CUDA(cudaStreamSynchronize(stream));
        int rc = check_timeout();
        return rc ? rc : check();
    }
    int set_state_full(int b, const void* data) override {
        if (bad_buffer(b)) return SWE_B200_EINVAL;
        SWE_CUDA(cudaMemcpyAsync(vol[cur[b]], data, sizeof(S) * plane * 3, cudaMemcpyHostToDevice, stream));
        SWE_CUDA(cudaStreamSynchronize(stream));
        return SWE_B200_OK;
    }
    int get_state_full(int b, void* data) override {
        if (bad_buffer(b)) return SWE_B200_EINVAL;
        SWE_CUDA(cudaMemcpyAsync(data, vol[cur[b]], sizeof(S) * plane * 3, cudaMemcpyDeviceToHost, stream));
        SWE_CUDA(cudaStreamSynchronize(stream));
        return check_timeout();
    }
    int pull_ctl() {
        SWE_CUDA(cudaMemcpyAsync(h_ctl, ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, stream));
        SWE_CUDA(cudaStreamSynchronize(stream));
        return check_timeout();      // multi-rank: a wait that gave up means halo rows / CFL maxima are stale
    }
    int set_time(double t, double dt) override {
        int rc = pull_ctl();
        if (rc) return rc;
        primed = false;
        h_ctl->t = (double)(S)t;
        h_ctl->dt = dt;
        h_ctl->dt_s = (double)(S)dt;
        h_ctl->dt_report = h_ctl->dt_s;
        SWE_CUDA(cudaMemcpyAsync(ctl, h_ctl, sizeof(Ctl), cudaMemcpyHostToDevice, stream));
        SWE_CUDA(cudaStreamSynchronize(stream));
        return SWE_B200_OK;
    }
    int get_time(double* t, double* dt) override {
        int rc = pull_ctl();
        if (rc) return rc;
        *t = h_ctl->t;
        *dt = h_ctl->dt_report;
        return SWE_B200_OK;
    }

    // ---- fine-grained operators -------------------------------------------------------------------
    void launch_bc(S* d) {
        const int nxp = Ny - 2 * gc, nyp = Nx - 2 * gc;
        k_bc<S><<<(nxp + 127) / 128, 128, 0, stream>>>(d, Nx, Ny, plane, gc, cfg.bc, 0, 1, 1);
        k_bc<S><<<(nyp + 127) / 128, 128, 0, stream>>>(d, Nx, Ny, plane, gc, cfg.bc, 1, par.ylo_phys, par.yhi_phys);
        launches += 2;
    }
    int update_bc(int b) override {
        if (bad_buffer(b)) return SWE_B200_EINVAL;
        launch_bc(vol[cur[b]]);
        return check();
    }
    void launch_reconstruct(const S* in, int dir) {
        const bool pure = cfg.eq_kind == 1;
        if (cfg.recon == SWE_B200_RECON_LINEAR) {
            if (pure) k_reconstruct<P, 1, true><<<grd2(), blk2(), 0, stream>>>(par, bathy(), in, left, right, dir);
            else k_reconstruct<P, 1, false><<<grd2(), blk2(), 0, stream>>>(par, bathy(), in, left, right, dir);
        } else {
            if (pure) k_reconstruct<P, 0, true><<<grd2(), blk2(), 0, stream>>>(par, bathy(), in, left, right, dir);
            else k_reconstruct<P, 0, false><<<grd2(), blk2(), 0, stream>>>(par, bathy(), in, left, right, dir);
        }
        ++launches;
    }
    int reconstruct(int dir, int in_buffer) override {
        if (bad_buffer(in_buffer) || dir < 0 || dir > 1) return SWE_B200_EINVAL;
        int rc = ensure_face_buffers();
        if (rc) return rc;
        launch_reconstruct(vol[cur[in_buffer]], dir);
        return check();
    }
    int get_faces(void* l, void* r) override {
        int rc = ensure_face_buffers();
        if (rc) return rc;
        SWE_CUDA(cudaMemcpyAsync(l, left, sizeof(S) * plane * 3, cudaMemcpyDeviceToHost, stream));
        SWE_CUDA(cudaMemcpyAsync(r, right, sizeof(S) * plane * 3, cudaMemcpyDeviceToHost, stream));
        SWE_CUDA(cudaStreamSynchronize(stream));
        return SWE_B200_OK;
    }
    int get_wavespeeds(void* w) override {
        int rc = ensure_face_buffers();
        if (rc) return rc;
        SWE_CUDA(cudaMemcpyAsync(w, ws, sizeof(S) * plane, cudaMemcpyDeviceToHost, stream));
        SWE_CUDA(cudaStreamSynchronize(stream));
        return SWE_B200_OK;
    }
    // asynchronous snapshot: state -> device snapshot buffer on the compute stream (so the following steps may
    // overwrite the volumes), then device -> host on a copy stream while the simulation continues
    S* snap = nullptr;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t snap_ready = nullptr, snap_done = nullptr;
    bool snap_pending = false;
    int snapshot_begin(void* w, void* hu, void* hv) override {
        if (!snap) {
            SWE_CUDA(cudaMalloc(&snap, sizeof(S) * plane * 3));
            SWE_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
            SWE_CUDA(cudaEventCreateWithFlags(&snap_ready, cudaEventDisableTiming));
            SWE_CUDA(cudaEventCreateWithFlags(&snap_done, cudaEventDisableTiming));
        }
        if (snap_pending) SWE_CUDA(cudaStreamWaitEvent(stream, snap_done, 0));     // the previous snapshot still reads `snap`
        SWE_CUDA(cudaMemcpyAsync(snap, vol[cur[0]], sizeof(S) * plane * 3, cudaMemcpyDeviceToDevice, stream));
        SWE_CUDA(cudaEventRecord(snap_ready, stream));
        SWE_CUDA(cudaStreamWaitEvent(copy_stream, snap_ready, 0));
        void* dst[3] = {w, hu, hv};
        for (int k = 0; k < 3; ++k)
            SWE_CUDA(cudaMemcpy2DAsync(dst[k], sizeof(S) * cfg.nx, snap + (size_t)k * plane + (size_t)gc * Nx + gc, sizeof(S) * Nx,
                                       sizeof(S) * cfg.nx, cfg.ny, cudaMemcpyDeviceToHost, copy_stream));
        SWE_CUDA(cudaEventRecord(snap_done, copy_stream));
        snap_pending = true;
        return SWE_B200_OK;
    }
    int snapshot_wait() override {
        if (!snap_pending) return SWE_B200_OK;
        SWE_CUDA(cudaEventSynchronize(snap_done));
        snap_pending = false;
        return SWE_B200_OK;
    }
    // scalar reductions over the interior of a buffer (TotalWaterVolume-style callbacks): 64 bytes come back
    double* diag_buf = nullptr;
    int diagnostics(int b, double out[8]) override {
        if (bad_buffer(b)) return SWE_B200_EINVAL;
        const int nb = 148 * 4;
        if (!diag_buf) SWE_CUDA(cudaMalloc(&diag_buf, sizeof(double) * DIAG_SLOTS * (nb + 1)));
        k_diag_partial<P><<<nb, 256, 0, stream>>>(vol[cur[b]], bathy(), cfg.nx, cfg.ny, Nx, gc, plane, cfg.depth_cutoff, diag_buf);
        k_diag_final<<<1, 256, 0, stream>>>(diag_buf, nb, diag_buf + (size_t)nb * DIAG_SLOTS);
        launches += 2;
        SWE_CUDA(cudaMemcpyAsync(out, diag_buf + (size_t)nb * DIAG_SLOTS, sizeof(double) * DIAG_SLOTS, cudaMemcpyDeviceToHost, stream));
        SWE_CUDA(cudaStreamSynchronize(stream));
        int rc = check_timeout();
        return rc ? rc : check();
    }
    // system.jl:62-75 on device buffers; the per-direction max lands in ctl->maxspeed_bits
    void launch_time_derivative(S* out, const S* state) {
        const bool pure = cfg.eq_kind == 1;
        for (int dir = 0; dir < 2; ++dir) {
            launch_reconstruct(state, dir);
            if (dir == 0) {
                if (pure) k_flux<P, true, 0><<<grd2(), blk2(), 0, stream>>>(par, out, left, right, ws, ctl);
                else k_flux<P, false, 0><<<grd2(), blk2(), 0, stream>>>(par, out, left, right, ws, ctl);
            } else {
                if (pure) k_flux<P, true, 1><<<grd2(), blk2(), 0, stream>>>(par, out, left, right, ws, ctl);
                else k_flux<P, false, 1><<<grd2(), blk2(), 0, stream>>>(par, out, left, right, ws, ctl);
            }
            ++launches;
            if (cfg.has_bottom_source && !pure) {
                k_bottom<P><<<grd2(), blk2(), 0, stream>>>(par, bathy(), out, left, right, dir);
                ++launches;
            }
        }
        for (int pass = 0; pass < 2; ++pass) {
            const bool do_infil = (pass == 0) == (cfg.infil_first != 0);
            if (do_infil && cfg.infil_kind != SWE_B200_INFIL_NONE) {
                k_point_source<P><<<grd1(plane), 256, 0, stream>>>(out, plane, ctl, 1, factor);
                ++launches;
            }
            if (!do_infil && cfg.rain_kind != SWE_B200_RAIN_NONE) {
                k_point_source<P><<<grd1(plane), 256, 0, stream>>>(out, plane, ctl, 0, rainp);
                ++launches;
            }
        }
    }
    int add_time_derivative(int in_buffer, int out_buffer, double t, double maxspeed[2]) override {
        if (bad_buffer(in_buffer) || bad_buffer(out_buffer)) return SWE_B200_EINVAL;
        int rc = ensure_face_buffers();
        if (rc) return rc;
        rc = pull_ctl();
        if (rc) return rc;
        primed = false;
        const double t_saved = h_ctl->t;
        h_ctl->t = (double)(S)t;
        SWE_CUDA(cudaMemcpyAsync(ctl, h_ctl, sizeof(Ctl), cudaMemcpyHostToDevice, stream));
        k_ctl_begin<<<1, 1, 0, stream>>>(ctl, rt, ip, 0.0, 0, 0, 0.0, 0.0);
        ++launches;
        launch_time_derivative(vol[cur[out_buffer]], vol[cur[in_buffer]]);
        rc = pull_ctl();
        if (rc) return rc;
        maxspeed[0] = (double)Bits<S>::dec(h_ctl->maxspeed_bits[0]);
        maxspeed[1] = (double)Bits<S>::dec(h_ctl->maxspeed_bits[1]);
        h_ctl->t = t_saved;
        SWE_CUDA(cudaMemcpyAsync(ctl, h_ctl, sizeof(Ctl), cudaMemcpyHostToDevice, stream));
        SWE_CUDA(cudaStreamSynchronize(stream));
        return check();
    }
    void launch_friction(S* out, const S* prev, double dt_host, int use_host) {
        if (cfg.friction == SWE_B200_FRICTION_NONE) return;
        if (cfg.eq_kind == 1) k_friction<P, true><<<grd2(), blk2(), 0, stream>>>(par, bathy(), out, prev, ctl, dt_host, use_host);
        else k_friction<P, false><<<grd2(), blk2(), 0, stream>>>(par, bathy(), out, prev, ctl, dt_host, use_host);
        ++launches;
    }
    int implicit_substep(int out_buffer, int prev_buffer, double dt) override {
        if (bad_buffer(out_buffer) || bad_buffer(prev_buffer)) return SWE_B200_EINVAL;
        launch_friction(vol[cur[out_buffer]], vol[cur[prev_buffer]], dt, 1);
        return check();
    }
    void launch_postproc(S* out) {
        if (cfg.eq_kind != 0) return;   // post_proc_substep!(::System, ::Equation) = nothing for Pure
        k_postproc<P><<<grd2(), blk2(), 0, stream>>>(par, bathy(), out);
        ++launches;
    }
    int post_proc_substep(int b) override {
        if (bad_buffer(b)) return SWE_B200_EINVAL;
        launch_postproc(vol[cur[b]]);
        return check();
    }

    // ---- one perform_step!, unfused: the reference's kernel sequence 1:1 (simulator.jl:145-177) ------
    void step_unfused(int advance) {
        primed = false;
        const int nsub = cfg.stepper == SWE_B200_RK2 ? 2 : 1;
        const long long n3 = plane * 3;
        for (int s = 0; s < nsub; ++s) {
            S* out = vol[cur[s + 1]];
            const S* state = vol[cur[s]];
            k_fill<S><<<grd1(n3), 256, 0, stream>>>(out, n3, (S)0);                    // rungekutta.jl:28-30
            ++launches;
            launch_time_derivative(out, state);
            if (s == 0) { k_ctl_dt<S><<<1, 1, 0, stream>>>(ctl, (double)par_dx(), (double)par_dy(), (S)cfg.cfl); ++launches; }
            k_axpy<P><<<grd1(n3), 256, 0, stream>>>(out, state, n3, ctl, s == 0 ? 0 : 1);   // :39-41
            ++launches;
            if (s == 1) { k_rkavg<P><<<grd1(n3), 256, 0, stream>>>(out, vol[cur[0]], n3); ++launches; }   // :44-49
            launch_friction(out, state, 0.0, 0);
            launch_postproc(out);
            launch_bc(out);
        }
        std::swap(cur[0], cur[nsub]);
        if (advance) { k_ctl_advance<S><<<1, 1, 0, stream>>>(ctl); ++launches; }
    }
    double par_dx() const { return (cfg.extent[1] - cfg.extent[0]) / cfg.nx; }
    double par_dy() const { return (cfg.extent[3] - cfg.extent[2]) / (cfg.global_ny > 0 ? cfg.global_ny : cfg.ny); }

    // ---- one perform_step!, fused --------------------------------------------------------------------
    template <int STAGE>
    void launch_stage(const S* U, const S* D, S* out) {
        StageArgs<P> a;
        a.p = par; a.U = U; a.D = D; a.out = out; a.B = B; a.wrapLo = wrap_ptr(0); a.wrapHi = wrap_ptr(1); a.Bp = Bp; a.factor = factor; a.rainp = rainp; a.ctl = ctl;
        a.rain_kind = cfg.rain_kind; a.infil_kind = cfg.infil_kind; a.infil_first = cfg.infil_first;
        // rows marched per block: every block pays a ~4.5-row prologue, and the launch runs in waves of 148 x 4
        // resident blocks, so the rows are chosen to minimise waves x (rows + 4.5).  Large grids end up near 64; a
        // 1024^2 grid gets 18 rows = one full wave (at 64 rows it would be 640 warps, one per scheduler:
        // measured 1024^2 Float64 K1+K2 per step: 64 rows 0.318, 32: 0.191, 16: 0.153, 8: 0.138 ms).
        const int nwarps = (cfg.nx + FUSED_VW - 1) / FUSED_VW;
        const long long nbx = (nwarps + FUSED_WPB - 1) / FUSED_WPB;
        a.rows = FUSED_ROWS;
        if (const char* e = std::getenv("SWE_B200_FUSED_ROWS")) a.rows = std::max(4, atoi(e));
        else {
            double best = 1e300;
            for (int r = 6; r <= FUSED_ROWS; ++r) {
                const long long blocks = nbx * ((cfg.ny + r - 1) / r);
                const double cost = (double)((blocks + 148 * 4 - 1) / (148 * 4)) * (r + 4.5);
                if (cost <= best) { best = cost; a.rows = r; }
            }
        }
        dim3 g((nwarps + FUSED_WPB - 1) / FUSED_WPB, (cfg.ny + a.rows - 1) / a.rows);
        dim3 b(FUSED_WPB * 32);
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (timing_on) {
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0, stream);
        }
        const bool lin = cfg.recon == SWE_B200_RECON_LINEAR;
        if (x2_ok) launch_stage_x2<STAGE>(U, D, out);
        else if (d1_ok) launch_stage_d1<STAGE>(U, D, out);
        else if (B) {
            if (lin) k_stage<P, 1, true, STAGE><<<g, b, 0, stream>>>(a);
            else k_stage<P, 0, true, STAGE><<<g, b, 0, stream>>>(a);
        } else {
            if (lin) k_stage<P, 1, false, STAGE><<<g, b, 0, stream>>>(a);
            else k_stage<P, 0, false, STAGE><<<g, b, 0, stream>>>(a);
        }
        ++launches;
        if (timing_on) {
            cudaEventRecord(e1, stream);
            ev.push_back(e0); ev.push_back(e1);
            ev_stage.push_back(STAGE);
        }
    }
    // ---- multi-GPU: peer pushes of halo rows / CFL maxima (kernels_exchange.cuh) -------------------------
    // kind 0: dU rows (stamp slots 0/1), kind 2: U rows (slots 2/3); rows_only=false also publishes the speeds
    PushArgs<S> push_args(const S* src, int vol_index, int kind, bool with_speed, bool rows, bool with_dt) {
        PushArgs<S> a{};
        a.src = src;
        a.src_plane = plane;
        a.Nx = Nx; a.Ny = Ny; a.gc = gc;
        a.rows = rows ? gc : 0;
        if (rows) a.stamp = ++xseq;
        if (lo_rank >= 0) {
            a.dst_lo = (S*)(peer_base[lo_rank] + (size_t)vol_index * peer_vol_bytes[lo_rank]);
            a.lo_plane = peer_plane[lo_rank];
            a.lo_Ny = (int)peer_Ny[lo_rank];
            a.flag_lo = &peer_mbox[lo_rank]->halo_stamp[kind + 1];   // I am the lower rank's UPPER neighbour
        }
        if (hi_rank >= 0) {
            a.dst_hi = (S*)(peer_base[hi_rank] + (size_t)vol_index * peer_vol_bytes[hi_rank]);
            a.hi_plane = peer_plane[hi_rank];
            a.flag_hi = &peer_mbox[hi_rank]->halo_stamp[kind + 0];
        }
        a.with_speed = with_speed ? 1 : 0;
        a.nranks = cfg.nranks; a.rank = cfg.rank; a.seq = sseq; a.ctl = ctl;
        for (int r = 0; r < cfg.nranks; ++r) a.mbox[r] = peer_mbox[r];
        a.ticket = ticket;
        a.ctl_rw = ctl;
        a.dt.enabled = with_dt ? 1 : 0;
        if (with_dt) {      // the last block of the push also gathers the maxima and computes dt (was k_ctl_dt_multi)
            a.dt.dx = par_dx(); a.dt.dy = par_dy(); a.dt.cfl = (double)(S)cfg.cfl;
            a.dt.nranks = cfg.nranks; a.dt.has_lo = recv_lo; a.dt.has_hi = recv_hi;
            a.dt.seq = sseq; a.dt.dU_stamp = rows ? a.stamp : 0ull;
        }
        return a;
    }
    int push_blocks(int rows) const {
        const long long per_side = (long long)rows * Nx;
        return (int)std::max<long long>(1, std::min<long long>((per_side / 4 + 255) / 256, 8));
    }
    // kind 0: dU rows (stamp slots 0/1), kind 2: U rows (slots 2/3); rows_only=false also publishes the speeds
    void push_halo(const S* src, int vol_index, int kind, bool with_speed, bool rows = true, bool with_dt = false) {
        const PushArgs<S> a = push_args(src, vol_index, kind, with_speed, rows, with_dt);
        k_halo_push<S><<<dim3(push_blocks(a.rows), a.rows > 0 ? 6 : 1), 256, 0, stream>>>(a);
        ++launches;
        if (rows && kind == 2) pending_U = a.stamp;
    }
    // the last launch of a fused step (k_tail): ghost fill + U-row push + t += dt + the next step's ctl_begin + U-row wait
    void launch_tail(S* next, int next_index, int advance, const BeginKey& key) {
        TailArgs<S> a{};
        a.d = next; a.Nx = Nx; a.Ny = Ny; a.gc = gc; a.bc = cfg.bc; a.ylo_phys = par.ylo_phys; a.yhi_phys = par.yhi_phys;
        a.advance = advance; a.plane = plane; a.ctl = ctl;
        a.begin.rt = rt; a.begin.ip = ip;
        a.begin.max_dt_host = key.max_dt_host; a.begin.t_end = key.t_end; a.begin.max_timestep = key.max_timestep;
        a.begin.use_t_end = key.use_t_end; a.begin.match_endtime = key.match; a.begin.enabled = 1;
        const long long nthreads = (long long)((Ny - 2 * gc) + (Nx - 2 * gc)) * gc * 6;
        a.nb_bc = (int)((nthreads + 255) / 256);
        a.multi = peers_open ? 1 : 0;
        int nblocks = a.nb_bc;
        if (peers_open) {
            a.push = push_args(next, next_index, 2, false, true, false);
            a.nbp = push_blocks(a.push.rows);
            a.has_lo = recv_lo; a.has_hi = recv_hi;
            a.mb = mbox;
            nblocks += 6 * a.nbp;
        }
        k_tail<S><<<nblocks, 256, 0, stream>>>(a);
        ++launches;
        primed = true;
        primed_key = key;
    }

    void step_fused(int advance, const BeginKey& key) {
        const bool multi = peers_open;
        S* U = vol[cur[0]];
        S* D = vol[cur[1]];
        mark("ctl_begin");
        launch_stage<1>(U, nullptr, D);
        mark("K1");
        const bool rk2 = cfg.stepper == SWE_B200_RK2;
        const bool neumann = cfg.bc == SWE_B200_BC_NEUMANN;
        const int nline = (int)((Ny + Nx - 1 + 127) / 128);
        if (neumann) {       // dU of the self-evolving outermost high ghost column / row = the per-cell sources (before the
                             // dU rows go to the neighbours: their halo rows carry the ghost column too)
            k_neumann_dU<P><<<nline, 128, 0, stream>>>(par, D, ctl, factor, rainp, cfg.rain_kind, cfg.infil_kind, cfg.infil_first);
            ++launches;
        }
        if (multi) {
            ++sseq;
            // dU rows (RK2 only) + this rank's CFL maxima; the same launch then gathers all maxima -> dt
            push_halo(D, cur[1], 0, true, rk2, true);
        } else {
            k_ctl_dt<S><<<1, 1, 0, stream>>>(ctl, par_dx(), par_dy(), (S)cfg.cfl);
            ++launches;
        }
        mark("dt(+push dU)");
        S* next;
        int next_index;
        if (rk2) {
            next = vol[cur[2]];
            next_index = cur[2];
            launch_stage<2>(U, D, next);
            mark("K2");
            if (neumann) { k_neumann_ghosts<P, true><<<nline, 128, 0, stream>>>(par, bathy(), U, D, next, ctl); ++launches; }
            std::swap(cur[0], cur[2]);
        } else {
            dim3 b(32, 8), g((cfg.nx + 31) / 32, (cfg.ny + 7) / 8);
            if (B) k_update_fe<P, true><<<g, b, 0, stream>>>(par, bathy(), U, D, ctl);
            else k_update_fe<P, false><<<g, b, 0, stream>>>(par, bathy(), U, D, ctl);
            ++launches;
            if (neumann) { k_neumann_ghosts<P, false><<<nline, 128, 0, stream>>>(par, bathy(), U, D, D, ctl); ++launches; }
            next = D;
            next_index = cur[1];
            std::swap(cur[0], cur[1]);
        }
        launch_tail(next, next_index, advance, key);
        mark("tail");
    }
    void one_step(double max_dt_host, int use_t_end, int match, double t_end, double max_timestep, int advance) {
        const BeginKey key{max_dt_host, t_end, max_timestep, use_t_end, match};
        if (peers_open && fused_ok && pending_U) {   // + wait for the neighbours' U rows of set_state_interior's exchange
            k_ctl_begin_multi<<<1, 1, 0, stream>>>(ctl, rt, ip, max_dt_host, use_t_end, match, t_end, max_timestep, mbox,
                                                   recv_lo, recv_hi, pending_U);
            pending_U = 0;
            ++launches;
        } else if (!(fused_ok && primed && primed_key == key)) {
            k_ctl_begin<<<1, 1, 0, stream>>>(ctl, rt, ip, max_dt_host, use_t_end, match, t_end, max_timestep);
            ++launches;
        }
        if (fused_ok) step_fused(advance, key);
        else step_unfused(advance);
    }
    int prepare() {
        if (!fused_ok) return ensure_face_buffers();
        for (int side = 0; side < 2; ++side)
            if (needs_wrap(side) && !wrap_set[side]) {
                err = "this strip's ghost rows are a wrap-around image of the far end of the domain: hand over the bathymetry of the "
                      "source rows with swe_b200_set_wrap_bathymetry first";
                return SWE_B200_EINVAL;
            }
        return SWE_B200_OK;
    }

    int step(double max_dt, int advance_time) override {
        int rc = prepare();
        if (rc) return rc;
        one_step(max_dt, 0, 0, 0.0, 0.0, advance_time);
        return check();
    }
    int run_steps(int64_t n, double maximum_timestep, swe_b200_timing* timing) override {
        int rc = prepare();
        if (rc) return rc;
        const double inf = std::numeric_limits<double>::infinity();
        cudaEvent_t t0 = nullptr, t1 = nullptr;
        const int64_t l0 = launches;
        if (timing) {
            SWE_CUDA(cudaEventCreate(&t0));
            SWE_CUDA(cudaEventCreate(&t1));
            timing_on = true;
            { const char* e = std::getenv("SWE_B200_TRACE"); trace_on = e && e[0] == '1'; }
            mark("start");
            SWE_CUDA(cudaStreamSynchronize(stream));
            if (peers_open) {      // align the GPUs: every rank's timed region starts when the last rank is ready
                BarrierArgs ba{};
                for (int r = 0; r < cfg.nranks; ++r) ba.mbox[r] = peer_mbox[r];
                ba.nranks = cfg.nranks; ba.rank = cfg.rank; ba.seq = ++bseq;
                k_rank_barrier<<<1, 32, 0, stream>>>(ba);
            }
            SWE_CUDA(cudaEventRecord(t0, stream));
        }
        for (int64_t s = 0; s < n; ++s) one_step(inf, 0, 0, 0.0, maximum_timestep, 1);
        if (timing) {
            SWE_CUDA(cudaEventRecord(t1, stream));
            SWE_CUDA(cudaStreamSynchronize(stream));
            timing_on = false;
            if (trace_on && tr_ev.size() > 1) {
                std::vector<std::pair<const char*, double>> acc2;
                for (size_t q = 1; q < tr_ev.size(); ++q) {
                    float m = 0;
                    cudaEventElapsedTime(&m, tr_ev[q - 1], tr_ev[q]);
                    bool found = false;
                    for (auto& pr : acc2) if (pr.first == tr_label[q]) { pr.second += m; found = true; }
                    if (!found) acc2.push_back({tr_label[q], (double)m});
                }
                fprintf(stderr, "[swe_b200 trace rank %d] per step (us):", cfg.rank);
                for (auto& pr : acc2) fprintf(stderr, " %s=%.1f", pr.first, 1e3 * pr.second / (double)n);
                fprintf(stderr, "\n");
                for (auto e : tr_ev) cudaEventDestroy(e);
                tr_ev.clear(); tr_label.clear();
            }
            float ms = 0;
            SWE_CUDA(cudaEventElapsedTime(&ms, t0, t1));
            timing->total_ms = ms;
            timing->launches = launches - l0;
            double acc[3] = {0, 0, 0};
            for (size_t q = 0; q + 1 < ev.size(); q += 2) {
                float m = 0;
                cudaEventElapsedTime(&m, ev[q], ev[q + 1]);
                acc[ev_stage[q / 2]] += m;
            }
            timing->stage1_ms = acc[1];
            timing->stage2_ms = acc[2];
            timing->stage_kernel_ms = acc[1] + acc[2];
            timing->stage_launches = (int64_t)(ev.size() / 2);
            for (auto e : ev) cudaEventDestroy(e);
            ev.clear();
            ev_stage.clear();
            cudaEventDestroy(t0);
            cudaEventDestroy(t1);
            int rc2 = check_timeout();
            if (rc2) return rc2;
        }
        return check();
    }
    // simulator.jl:179-212.  dt and t live on the device; the host reads t back every `batch` steps.
    // Steps issued after t reached t_end are exact no-ops (dt = 0) and are not counted.
    int simulate_to_time(double t_end, int match_endtime, double maximum_timestep, int64_t* nsteps) override {
        int rc = prepare();
        if (rc) return rc;
        rc = pull_ctl();
        if (rc) return rc;
        const long long n0 = h_ctl->nsteps;
        int batch = 1, stalled = 0;
        int result = SWE_B200_OK;
        while (h_ctl->t < t_end) {
            const double t_before = h_ctl->t;
            for (int s = 0; s < batch; ++s) one_step(0.0, 1, match_endtime, t_end, maximum_timestep, 1);
            rc = check();
            if (rc) return rc;
            rc = pull_ctl();
            if (rc) return rc;
            // the reference would loop forever on these (simulator.jl:199-210); an ABI call must come back
            if (std::isnan(h_ctl->t) || std::isnan(h_ctl->dt)) { err = "simulate_to_time: NaN time step (the state blew up)"; result = SWE_B200_EDOMAIN; break; }
            stalled = h_ctl->t == t_before ? stalled + 1 : 0;
            if (stalled >= 3) {      // >= 1 + 2 + 4 steps without any progress in t: dt == 0 (e.g. an infinite wave speed)
                err = "simulate_to_time: the time step is 0, t does not advance (infinite wave speed?)";
                result = SWE_B200_EDOMAIN;
                break;
            }
            if (batch < 32) batch *= 2;
        }
        if (nsteps) *nsteps = h_ctl->nsteps - n0;
        return result;
    }

    int enumerate(int kind, int dir, int g, int64_t* out, int64_t* count) override {
        long long na, nb;
        int per;
        if (kind == 0) { na = Nx; nb = Ny; per = 2; }
        else if (kind == 1) { na = dir == 0 ? Nx - 2 * g : Nx - 2 * gc; nb = dir == 0 ? Ny - 2 * gc : Ny - 2 * g; per = 6; }
        else { na = dir == 0 ? gc : Nx - 2 * gc; nb = dir == 0 ? Ny - 2 * gc : gc; per = 2; }
        if (na < 0 || nb < 0) { err = "empty index space"; return SWE_B200_EINVAL; }
        *count = na * nb;
        if (!out || *count == 0) return SWE_B200_OK;
        long long* d = nullptr;
        SWE_CUDA(cudaMalloc(&d, sizeof(long long) * per * (*count)));
        dim3 b(32, 8), gr((unsigned)((na + 31) / 32), (unsigned)((nb + 7) / 8));
        k_enumerate<<<gr, b, 0, stream>>>(kind, dir, g, Nx, Ny, gc, d);
        ++launches;
        SWE_CUDA(cudaMemcpyAsync(out, d, sizeof(long long) * per * (*count), cudaMemcpyDeviceToHost, stream));
        SWE_CUDA(cudaStreamSynchronize(stream));
        cudaFree(d);
        return check();
    }

    int get_ipc_handle(void* out128) override {
        IpcBlob b;
        std::memset(&b, 0, sizeof(b));
        SWE_CUDA(cudaIpcGetMemHandle(&b.handle, slab));
        b.pid = (long long)getpid();
        b.base = slab;
        b.Ny = Ny; b.plane = plane; b.vol_bytes = (long long)vol_bytes; b.mbox_offset = (long long)mbox_offset;
        b.device = cfg.device; b.rank = cfg.rank;
        std::memcpy(out128, &b, sizeof(b));
        return SWE_B200_OK;
    }
    int open_peers(const void* all128) override {
        if (cfg.nranks <= 1) { err = "open_peers on a single-rank handle"; return SWE_B200_EINVAL; }
        const IpcBlob* blobs = (const IpcBlob*)all128;
        SWE_CUDA(cudaSetDevice(cfg.device));
        for (int r = 0; r < cfg.nranks; ++r) {
            const IpcBlob& b = blobs[r];
            if (b.rank != r) { err = "peer handles must be ordered by rank"; return SWE_B200_EINVAL; }
            if (r == cfg.rank) {
                peer_base[r] = slab;
            } else if (b.pid == (long long)getpid()) {      // same process: plain peer access
                int can = 0;
                SWE_CUDA(cudaDeviceCanAccessPeer(&can, cfg.device, b.device));
                if (!can && b.device != cfg.device) { err = "no peer access between the devices"; return SWE_B200_ECUDA; }
                if (b.device != cfg.device) {
                    cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
                        err = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e);
                        return SWE_B200_ECUDA;
                    }
                    cudaGetLastError();
                }
                peer_base[r] = (char*)b.base;
            } else {
                void* ptr = nullptr;
                SWE_CUDA(cudaIpcOpenMemHandle(&ptr, b.handle, cudaIpcMemLazyEnablePeerAccess));
                peer_base[r] = (char*)ptr;
                peer_ipc[r] = true;
            }
            peer_Ny[r] = b.Ny; peer_plane[r] = b.plane; peer_vol_bytes[r] = b.vol_bytes;
            peer_mbox[r] = (Mailbox*)(peer_base[r] + b.mbox_offset);
        }
        peers_open = true;
        return SWE_B200_OK;
    }
    int close_peers() override {
        if (!peers_open) return SWE_B200_OK;
        SWE_CUDA(cudaStreamSynchronize(stream));
        int rc = check_timeout();
        for (int r = 0; r < MAX_RANKS; ++r) {
            if (peer_ipc[r] && peer_base[r]) cudaIpcCloseMemHandle(peer_base[r]);
            peer_ipc[r] = false;
            peer_base[r] = nullptr;
            peer_mbox[r] = nullptr;
        }
        peers_open = false;
        return rc;
    }
    int check_timeout() {
        if (!peers_open) return SWE_B200_OK;
        unsigned long long t = 0;
        SWE_CUDA(cudaMemcpyAsync(&t, &mbox->timeout, sizeof(t), cudaMemcpyDeviceToHost, stream));
        SWE_CUDA(cudaStreamSynchronize(stream));
        if (t) {
            err = "multi-GPU exchange timed out waiting for a neighbour: halo rows / CFL maxima are stale, the state of "
                  "this handle is invalid";
            return SWE_B200_ECUDA;
        }
        return SWE_B200_OK;
    }
    // which kernel family serves step / run_steps / simulate_to_time on this handle (swe_b200_kernel_family)
    int kernel_family() const override {
        if (!fused_ok) return SWE_B200_FAMILY_PER_OPERATOR;
        if (x2_ok) return SWE_B200_FAMILY_FUSED_X2;
        if (d1_ok) return SWE_B200_FAMILY_FUSED_D1;
        return SWE_B200_FAMILY_FUSED_GENERIC;
    }
    const void* state_ptr(int b) const override { return vol[cur[b]]; }
    int elem_bytes() const override { return (int)sizeof(S); }
    void dims(int out[4]) const override { out[0] = cfg.nx; out[1] = cfg.ny; out[2] = gc; out[3] = cfg.dtype; }
    // per-cell comparison of this handle's state with another handle's (same grid, same dtype, same device) on the
    // device: parity checks at sizes where a D2H copy + numpy diff of both states would dominate the test
    int compare(const void* other_state, int b, const double* thresholds, double out[SWE_B200_COMPARE_SLOTS]) override {
        static_assert(CMP_SLOTS == SWE_B200_COMPARE_SLOTS, "compare slots");
        CmpThr thr{};
        if (thresholds) { thr.on = 1; for (int k = 0; k < 3; ++k) thr.v[k] = thresholds[k]; }
        if (bad_buffer(b)) return SWE_B200_EINVAL;
        const int nb = 148 * 4;
        double* buf = nullptr;
        SWE_CUDA(cudaMalloc(&buf, sizeof(double) * CMP_SLOTS * (nb + 1)));
        k_cmp_partial<P><<<nb, 256, 0, stream>>>(vol[cur[b]], (const S*)other_state, bathy(), cfg.nx, cfg.ny, Nx, gc, plane,
                                                 cfg.depth_cutoff, thr, buf);
        k_cmp_final<<<1, 256, 0, stream>>>(buf, nb, buf + (size_t)nb * CMP_SLOTS);
        launches += 2;
        SWE_CUDA(cudaMemcpyAsync(out, buf + (size_t)nb * CMP_SLOTS, sizeof(double) * CMP_SLOTS, cudaMemcpyDeviceToHost, stream));
        SWE_CUDA(cudaStreamSynchronize(stream));
        cudaFree(buf);
        return check();
    }
    int64_t launch_count() const override { return launches; }
    int device() const override { return cfg.device; }
};

template <class P>
Engine* make_engine_t(const swe_b200_config& cfg, std::string& err) {
    auto* e = new EngineT<P>();
    int rc = e->init(cfg);
    if (rc != SWE_B200_OK) {
        err = e->err;
        delete e;
        return nullptr;
    }
    return e;
}

SWE_NS_END  // namespace swe

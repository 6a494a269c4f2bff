// kernels_exchange.cuh -- multi-GPU y-strip exchange over NVLink peer memory (one process per GPU).
//
// The reference is single-device; the decomposition is new.  Per RK2 step each rank
//   * pushes the 2 outermost owned rows of dU (after K1) and of U^{n+1} (after K2 + BC) straight into
//     the ghost rows of its y-neighbours' buffers with peer stores (no staging, no host involvement),
//   * all-gathers the two directional CFL maxima through 8-byte mailbox slots in every peer,
// and the consumers (k_ctl_dt / k_ctl_begin) spin on monotonically increasing stamps written after a
// system-scope fence.  Every wait has a time-out so a dead peer cannot hang the GPU.
#pragma once
#include "kernels_unfused.cuh"

SWE_NS_BEGIN

constexpr int MAX_RANKS = 8;

struct Mailbox {
    unsigned long long speed[2][MAX_RANKS][2];     // [parity][rank][dir] bit patterns of the local maxima
    unsigned long long speed_stamp[2][MAX_RANKS];  // step sequence number the slot belongs to
    unsigned long long halo_stamp[4];              // 0: dU from lower, 1: dU from upper, 2: U from lower, 3: U from upper
    unsigned long long timeout;                    // != 0: a wait gave up (reported by the host as an error)
    unsigned long long arrive[MAX_RANKS];          // device-side rendezvous (k_rank_barrier): sequence number per rank
};

__device__ __forceinline__ bool spin_until(volatile unsigned long long* p, unsigned long long want, Mailbox* mb) {
    const long long t0 = clock64();
    if (*(volatile unsigned long long*)&mb->timeout) return false;     // an earlier wait gave up: do not spin again
    while (*p < want) {
        __nanosleep(64);
        if (clock64() - t0 > 6000000000ll) {   // ~3 s at 2 GHz
            mb->timeout = 1;
            return false;
        }
    }
    return true;
}

// wait for the neighbours' rows of kind `base` (0: dU, 2: U) to have arrived
__device__ __forceinline__ void wait_halo(Mailbox* mb, int base, int has_lo, int has_hi, unsigned long long stamp) {
    if (has_lo) spin_until(&mb->halo_stamp[base + 0], stamp, mb);
    if (has_hi) spin_until(&mb->halo_stamp[base + 1], stamp, mb);
    __threadfence_system();
}

static __global__ void k_wait_halo(Mailbox* mb, int base, int has_lo, int has_hi, unsigned long long stamp) {
    wait_halo(mb, base, has_lo, has_hi, stamp);
}
// start of a multi-GPU step in one launch: k_ctl_begin + the wait for the neighbours' U rows of the last exchange
static __global__ void k_ctl_begin_multi(Ctl* c, RainTable r, InfilPar f, double max_dt_host, int use_t_end,
                                         int match_endtime, double t_end, double max_timestep, Mailbox* mb, int has_lo,
                                         int has_hi, unsigned long long U_stamp) {
    ctl_begin(c, r, f, max_dt_host, use_t_end, match_endtime, t_end, max_timestep);
    if (U_stamp) wait_halo(mb, 2, has_lo, has_hi, U_stamp);
}

// multi-rank timestep: gather all ranks' maxima from the mailbox, then timestep_computer as in k_ctl_dt,
// then wait for the dU halo rows stage 2 needs.
struct DtArgs {
    double dx, dy, cfl;           // cfl already rounded to the realtype
    int enabled, nranks, has_lo, has_hi;
    unsigned long long seq, dU_stamp;
};
// Called by the first warp of the last block (all 32 lanes): lane r waits for rank r's stamp and reads its two maxima -- the
// 3 x nranks dependent loads of a one-thread loop (each an uncached read of peer-written memory) collapse into 3 -- then a
// warp max; lane 0 finishes the time step.  The maximum is order-independent, so every rank still computes identical bits.
template <class S>
__device__ __forceinline__ void ctl_dt_multi(Ctl* c, Mailbox* mb, const DtArgs& d) {
    const int par = (int)(d.seq & 1ull);
    const int lane = threadIdx.x & 31;
    unsigned long long bx = 0ull, by = 0ull;
    if (lane < d.nranks) {
        spin_until(&mb->speed_stamp[par][lane], d.seq, mb);
        bx = *(volatile unsigned long long*)&mb->speed[par][lane][0];
        by = *(volatile unsigned long long*)&mb->speed[par][lane][1];
    }
    if (d.dU_stamp) {        // the two halo stamps are polled by two more lanes meanwhile
        if (lane == 30 && d.has_lo) spin_until(&mb->halo_stamp[0], d.dU_stamp, mb);
        if (lane == 31 && d.has_hi) spin_until(&mb->halo_stamp[1], d.dU_stamp, mb);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long tx = __shfl_xor_sync(0xffffffffu, bx, o), ty = __shfl_xor_sync(0xffffffffu, by, o);
        bx = tx > bx ? tx : bx;
        by = ty > by ? ty : by;
    }
    if (lane != 0) return;
    c->maxspeed_bits[0] = bx;
    c->maxspeed_bits[1] = by;
    const S ax = Bits<S>::dec(bx);
    const S ay = Bits<S>::dec(by);
    double ddt = fmin(d.dx / ax, d.dy / ay);
    double max_dt = c->max_dt_host;
    if (c->use_t_end) max_dt = c->match_endtime ? (c->t_end - c->t) : (double)INFINITY;
    if (c->max_timestep > 0) max_dt = fmin(max_dt, c->max_timestep);
    double dt = fmin((S)d.cfl * ddt, max_dt);
    const bool past_end = c->use_t_end && !(c->t < c->t_end);
    if (past_end) dt = 0.0;
    c->dt = dt;
    c->dt_s = (double)(S)dt;
    if (!past_end) c->dt_report = c->dt_s;
    __threadfence_system();
}

template <class S>
struct PushArgs {
    const S* src;            // local volume (3 planes)
    S* dst_lo;               // lower neighbour's volume (peer pointer) or nullptr
    S* dst_hi;               // upper neighbour's volume (peer pointer) or nullptr
    long long src_plane, lo_plane, hi_plane;
    int Nx, Ny, gc, lo_Ny;   // lo_Ny: total rows of the lower neighbour's strip
    int rows;                // 0: no halo rows (speeds only)
    unsigned long long* flag_lo;   // in the lower neighbour's mailbox: "from upper" stamp
    unsigned long long* flag_hi;   // in the upper neighbour's mailbox: "from lower" stamp
    unsigned long long stamp;
    int with_speed, nranks, rank;
    unsigned long long seq;
    const Ctl* ctl;
    Mailbox* mbox[MAX_RANKS];      // every rank's mailbox (own included)
    unsigned int* ticket;          // local counter for last-block detection
    Ctl* ctl_rw;                   // for the fused timestep
    DtArgs dt;                     // dt.enabled: the last block also computes the step's dt (was a launch of its own)
};

// one (side, plane) run of halo rows: blockIdx-like coordinates (bx of nbx blocks, by = side * 3 + plane).  Each
// (side, plane) is ONE contiguous run of rows*Nx elements on both ends (the gc boundary rows are adjacent in memory),
// copied with 16-byte peer stores when the runs are aligned.
template <class S>
__device__ __forceinline__ void push_rows(const PushArgs<S>& a, int bx, int nbx, int by) {
    const long long per_side = (long long)a.rows * a.Nx;
    if (per_side <= 0) return;
    const int side = by / 3, k = by - 3 * side;
    const S* src;
    S* dst;
    if (side == 0) {   // my lowest owned rows -> lower neighbour's top ghost rows
        src = a.src + k * a.src_plane + (long long)a.gc * a.Nx;
        dst = a.dst_lo ? a.dst_lo + k * a.lo_plane + (long long)(a.lo_Ny - a.gc) * a.Nx : nullptr;
    } else {           // my highest owned rows -> upper neighbour's bottom ghost rows
        src = a.src + k * a.src_plane + (long long)(a.Ny - 2 * a.gc) * a.Nx;
        dst = a.dst_hi ? a.dst_hi + k * a.hi_plane : nullptr;
    }
    if (!dst) return;
    const long long t = bx * (long long)blockDim.x + threadIdx.x, nt = (long long)nbx * blockDim.x;
    constexpr int V = 16 / sizeof(S);
    if ((((unsigned long long)src | (unsigned long long)dst) & 15ull) == 0 && per_side % V == 0) {
        const int4* s4 = reinterpret_cast<const int4*>(src);
        int4* d4 = reinterpret_cast<int4*>(dst);
        for (long long q = t; q < per_side / V; q += nt) d4[q] = s4[q];
    } else {
        for (long long q = t; q < per_side; q += nt) dst[q] = src[q];
    }
}

template <class S>
__global__ void k_halo_push(const PushArgs<S> a) {
    // The strip's CFL maxima are final when this kernel starts (stage 1 is complete): publish them before the
    // row copies so that the all-gather latency overlaps them on every rank.
    if (a.with_speed && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x < a.nranks) {
        Mailbox* m = a.mbox[threadIdx.x];
        const int par = (int)(a.seq & 1ull);
        m->speed[par][a.rank][0] = a.ctl->maxspeed_bits[0];
        m->speed[par][a.rank][1] = a.ctl->maxspeed_bits[1];
        __threadfence_system();
        *(volatile unsigned long long*)&m->speed_stamp[par][a.rank] = a.seq;
    }
    push_rows(a, blockIdx.x, gridDim.x, blockIdx.y);
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = atomicAdd(a.ticket, 1u) == gridDim.x * gridDim.y - 1;
    __syncthreads();
    if (!last) return;
    if (threadIdx.x == 0) *a.ticket = 0;
    __threadfence_system();
    if (a.rows > 0) {
        if (threadIdx.x == 0 && a.flag_lo) *(volatile unsigned long long*)a.flag_lo = a.stamp;
        if (threadIdx.x == 1 && a.flag_hi) *(volatile unsigned long long*)a.flag_hi = a.stamp;
    }
    if (a.dt.enabled) {
        __syncthreads();
        if (threadIdx.x < 32) ctl_dt_multi<S>(a.ctl_rw, a.mbox[a.rank], a.dt);
    }
}

// ---- the tail of a fused step in ONE launch (was: push U rows, k_bc_advance, next step's k_ctl_begin + U-row wait):
//   blocks [0, nb_bc)            ghost fill of the new state (bc.jl:29-40, 81-110), both directions
//   blocks [nb_bc, nb_bc + 6 nbp) push of the owned boundary rows of U^{n+1} into the neighbours' ghost rows
//   block 0, thread 0             t += dt (simulator.jl:205), then the NEXT step's control set-up (ctl_begin)
//   last block to finish          raises the neighbours' stamps, then waits for their U rows to have arrived here
// The push need not wait for the ghost fill: no consumer reads the x-ghost columns of a halo row from memory (both
// stage kernels remap them).
struct BeginArgs {
    RainTable rt;
    InfilPar ip;
    double max_dt_host, t_end, max_timestep;
    int use_t_end, match_endtime, enabled;
};
template <class S>
struct TailArgs {
    S* d;
    int Nx, Ny, gc, bc, ylo_phys, yhi_phys, advance;
    long long plane;
    Ctl* ctl;
    BeginArgs begin;
    int multi, nb_bc, nbp, has_lo, has_hi;
    Mailbox* mb;
    PushArgs<S> push;
};
template <class S>
__device__ __forceinline__ void bc_fill(S* d, int Nx, int Ny, long long plane, int gc, int bc, int ylo_phys, int yhi_phys, long long t) {
    // one thread per (passive index, ghost layer, side, variable): a single load + store each, so the launch is
    // a few microseconds although the x ghosts are strided by a whole row
    const int nxp = Ny - 2 * gc;   // passive count for x ghosts
    const int nyp = Nx - 2 * gc;
    const int per = gc * 2 * 3;
    if (t >= (long long)(nxp + nyp) * per) return;
    const int line = (int)(t / per), r = (int)(t - (long long)line * per);
    const int k = r % 3, side = (r / 3) & 1, g = r / 6;
    const int dir = line < nxp ? 0 : 1;
    const int pp = (dir == 0 ? line : line - nxp) + gc;
    const int N = dir == 0 ? Nx : Ny;
    const size_t sa = dir == 0 ? 1 : (size_t)Nx, sp = dir == 0 ? (size_t)Nx : 1;
    const int normal = dir == 0 ? 1 : 2;
    const bool phys = dir == 0 ? true : (side == 0 ? ylo_phys != 0 : yhi_phys != 0);
    if (!phys) return;
    S* base = d + (size_t)k * plane + (size_t)pp * sp;
    const bool flip = bc == SWE_B200_BC_WALL && k == normal;
    const int a = side == 0 ? g : N - 1 - g;
    bool f;
    const S v = base[(size_t)bc_source(a, N, gc, bc, true, true, f, dir) * sa];
    base[(size_t)a * sa] = flip ? -v : v;          // a copy or a sign flip: exact for subnormals too
}
template <class S>
__device__ __forceinline__ void ctl_advance_begin(Ctl* ctl, int advance, const BeginArgs& b) {
    if (advance) {
        const bool skip = ctl->use_t_end && !(ctl->t < ctl->t_end);      // a step at t == t_end is an exact no-op
        if (!skip) {
            S tt = (S)ctl->t;
            tt += (S)ctl->dt_s;
            ctl->t = (double)tt;
            ctl->nsteps += 1;
        }
    }
    if (b.enabled) ctl_begin(ctl, b.rt, b.ip, b.max_dt_host, b.use_t_end, b.match_endtime, b.t_end, b.max_timestep);
}
template <class S>
__global__ void k_tail(const TailArgs<S> a) {
    const int b = blockIdx.x;
    if (b < a.nb_bc) {
        bc_fill<S>(a.d, a.Nx, a.Ny, a.plane, a.gc, a.bc, a.ylo_phys, a.yhi_phys, b * (long long)blockDim.x + threadIdx.x);
        if (b == 0 && threadIdx.x == 0) ctl_advance_begin<S>(a.ctl, a.advance, a.begin);
    } else {
        const int q = b - a.nb_bc;
        push_rows(a.push, q % a.nbp, a.nbp, q / a.nbp);
    }
    if (!a.multi) return;
    __threadfence_system();
    __syncthreads();
    __shared__ bool last;
    if (threadIdx.x == 0) last = atomicAdd(a.push.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    if (threadIdx.x == 0) *a.push.ticket = 0;
    __threadfence_system();
    if (threadIdx.x == 0 && a.push.flag_lo) *(volatile unsigned long long*)a.push.flag_lo = a.push.stamp;
    if (threadIdx.x == 1 && a.push.flag_hi) *(volatile unsigned long long*)a.push.flag_hi = a.push.stamp;
    __syncthreads();
    // the next stage-1 kernel reads the neighbours' rows: wait for them here, at the end of the step's last launch
    if (threadIdx.x == 0 && a.has_lo) spin_until(&a.mb->halo_stamp[2], a.push.stamp, a.mb);
    if (threadIdx.x == 1 && a.has_hi) spin_until(&a.mb->halo_stamp[3], a.push.stamp, a.mb);
    __threadfence_system();
}

// ---- device-side rendezvous of all ranks on the handles' streams: every rank announces sequence number `seq` in
// every peer's mailbox and waits until all have.  Used to align the GPUs before a timed region (host-side barriers
// release the ranks up to a millisecond apart, which is 10 % of a 20-step region at N = 8).
struct BarrierArgs {
    Mailbox* mbox[MAX_RANKS];
    int nranks, rank;
    unsigned long long seq;
};
static __global__ void k_rank_barrier(const BarrierArgs a) {
    if (threadIdx.x < a.nranks) {
        __threadfence_system();
        *(volatile unsigned long long*)&a.mbox[threadIdx.x]->arrive[a.rank] = a.seq;
        spin_until(&a.mbox[a.rank]->arrive[threadIdx.x], a.seq, a.mbox[a.rank]);
    }
    __threadfence_system();
}

SWE_NS_END  // namespace swe

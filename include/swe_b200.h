/* swe_b200.h -- C ABI of the B200-native backend for the 2D shallow-water finite-volume step of
 * sintefmath/VolumeFluxes.jl (module SinFVM).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.  A Julia
 * `ccall` shim (volumefluxes.jl_b200/julia/B200Backend.jl) or the Python/ctypes
 * host (volumefluxes.jl_b200/api.py) binds exactly these symbols.  Each entry point names the
 * reference interface it replaces (paths relative to the reference's src/).
 *
 * Conventions
 *  - every function returns 0 on success, <0 = SWE_B200_E*; swe_b200_last_error() gives text;
 *    nothing throws or aborts across the ABI.
 *  - host arrays are caller-owned, column-major exactly like the reference (`data[i,j,v]`,
 *    x fastest; volume/volume.jl, backends/buffer.jl:29-42) and have the element type selected by
 *    cfg.dtype; they are copied before the call returns.  The library owns all device memory.
 *  - one host thread per handle; work is stream-ordered on the handle's stream; only get_*,
 *    sync and simulate_to_time block.
 *  - "buffer" arguments index the simulator's substep_outputs (simulator.jl:64,86):
 *    0 = current state, 1..nsub = stage outputs.
 */
#ifndef SWE_B200_H
#define SWE_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWE_B200_ABI_VERSION 1

enum {
    SWE_B200_OK = 0,
    SWE_B200_EINVAL = -1,      /* ArgumentError / @assert in the reference            */
    SWE_B200_EDOMAIN = -2,     /* DomainError (bad bathymetry / factor size)           */
    SWE_B200_ECUDA = -3,       /* CUDA runtime failure                                 */
    SWE_B200_EUNSUPPORTED = -4 /* valid in the reference, not implemented on this path */
};

enum { SWE_B200_F32 = 0, SWE_B200_F64 = 1 };
enum { SWE_B200_BC_PERIODIC = 0, SWE_B200_BC_WALL = 1, SWE_B200_BC_NEUMANN = 2 }; /* grid.jl:21-28, bc.jl */
enum { SWE_B200_RECON_NONE = 0, SWE_B200_RECON_LINEAR = 1 };  /* reconstruction/ *.jl */
enum { SWE_B200_FE = 0, SWE_B200_RK2 = 1 };                   /* timestepper/ *.jl */
enum { SWE_B200_FRICTION_NONE = 0, SWE_B200_FRICTION_BH2021 = 1, SWE_B200_FRICTION_BSA2012 = 2,
       SWE_B200_FRICTION_FCG2016 = 3 };                       /* friction.jl:30-43    */
enum { SWE_B200_RAIN_NONE = 0, SWE_B200_RAIN_CONSTANT = 1, SWE_B200_RAIN_TIME_DEPENDENT = 2,
       SWE_B200_RAIN_PLANE = 3 /* FunctionalRain: per-cell values, see swe_b200_set_rain_plane */ };
enum { SWE_B200_INFIL_NONE = 0, SWE_B200_INFIL_CONSTANT = 1, SWE_B200_INFIL_HORTON = 2 };

/* Plain-old-data description of Grid + Equation + Reconstruction + System + Simulator
 * (grid.jl:82-90, equations/swe_2D.jl:22-29, reconstruction/linearreconstruction.jl:22-25,
 * system.jl:37-53, friction.jl:23-27, simulator.jl:72-91). */
typedef struct swe_b200_config {
    int32_t abi_version;       /* SWE_B200_ABI_VERSION                                         */
    int32_t dtype;             /* backend.realtype: SWE_B200_F32 / F64                          */
    int32_t nx, ny, gc;        /* interior cells of THIS handle's strip, ghost cells            */
    int32_t bc;                /* SWE_B200_BC_*                                                 */
    int32_t eq_kind;           /* 0 ShallowWaterEquations, 1 ShallowWaterEquationsPure          */
    int32_t recon;             /* SWE_B200_RECON_*                                              */
    int32_t stepper;           /* SWE_B200_FE / RK2                                             */
    int32_t friction;          /* SWE_B200_FRICTION_* (ImplicitFriction)                        */
    int32_t has_bathymetry;    /* 1: BottomTopography2D via set_bathymetry, 0: Constant(const_B)*/
    int32_t has_bottom_source; /* SourceTermBottom in the source-term list                      */
    int32_t rain_kind;         /* SWE_B200_RAIN_*                                               */
    int32_t infil_kind;        /* SWE_B200_INFIL_*                                              */
    int32_t infil_first;       /* 1: infiltration precedes rain in the source list (examples)   */
    int32_t n_rain;            /* TimeDependentRain table length (<=16)                         */
    int32_t strict;            /* 1: bit-faithful arithmetic (no FMA; f32 = Julia's mixed
                                  Float32/Float64 promotion), 0: fast native arithmetic         */
    int32_t fused;             /* 1: fused sm_100a stage kernels for step/simulate (default),
                                  0: one kernel per reference operator                           */
    int32_t device;            /* CUDA device ordinal                                            */
    /* y-strip decomposition (multi-GPU): this handle owns global rows [y_offset, y_offset+ny) */
    int32_t rank, nranks, global_ny, y_offset;
    int32_t reserved[3];
    double extent[4];          /* GLOBAL extent [x0 x1 y0 y1] (grid.jl:82); dx=(x1-x0)/nx,
                                  dy=(y1-y0)/global_ny                                           */
    double theta, rho, g, depth_cutoff, kappa, cfl, Cz, const_B, t0;
    double rain_rate;          /* ConstantRain rate per hour (divided by 3600 like the reference)*/
    double infil_const, f0, fc, k;
    double rain_rates[16], rain_times[16];
} swe_b200_config;

typedef struct swe_b200_handle swe_b200_handle;

/* Simulator(...) + ConservedSystem(...) + create_buffer (simulator.jl:72-91, system.jl:37-53,
 * backends/buffer.jl:29-42).  Validates like is_compatible (equation.jl:70-84, system.jl:81-83)
 * and `@assert gc > 1` for LinearReconstruction (linearreconstruction.jl:59). */
int swe_b200_create(const swe_b200_config* cfg, swe_b200_handle** out);
int swe_b200_destroy(swe_b200_handle* h);
const char* swe_b200_last_error(swe_b200_handle* h);   /* h may be NULL: last create() error */
int swe_b200_sync(swe_b200_handle* h);

/* BottomTopography2D(B, backend, grid) (bottom_topography.jl:51-59): corner values,
 * (nx+2gc+1) x (ny+2gc+1).  n_elems is checked like validate_bottom_topography (:78-82). */
int swe_b200_set_bathymetry(swe_b200_handle* h, const void* B, int64_t n_elems);
/* HortonInfiltration(...; factor) (sourceterms/source_term_infiltration.jl:37-46,61-65):
 * (nx+2gc) x (ny+2gc) doubles. */
int swe_b200_set_infiltration_factor(swe_b200_handle* h, const double* factor, int64_t n_elems);
/* FunctionalRain(rain_function, grid) (sourceterms/source_term_rain.jl:57-68): the reference calls an arbitrary
 * closure rain_function(t, x, y) per cell and adds the result to out.h (:73-75).  A precompiled kernel cannot run
 * the closure: with cfg.rain_kind = SWE_B200_RAIN_PLANE the host evaluates it at every cell centre
 * (cell_center, grid.jl:117-121, ALL (nx+2gc) x (ny+2gc) cells like for_each_cell) and hands the values over here --
 * once for a time-independent function, before every step for a time-dependent one (stream-ordered, may be called
 * between steps).  Units: what the closure returns (m/s; no /3600).  Zero until first set. */
int swe_b200_set_rain_plane(swe_b200_handle* h, const double* values, int64_t n_elems);

/* set_current_state! (simulator.jl:119-139): nx x ny per variable, then update_bc!. */
int swe_b200_set_state_interior(swe_b200_handle* h, const void* w, const void* hu, const void* hv);
/* collect(current_interior_state(sim).h/.hu/.hv) (volume/interior_volume_variable.jl:168-176) */
int swe_b200_get_state_interior(swe_b200_handle* h, void* w, void* hu, void* hv);
/* raw Volume access, (nx+2gc, ny+2gc, 3) (volume/volume.jl:24-71) */
int swe_b200_set_state_full(swe_b200_handle* h, int buffer, const void* data);
int swe_b200_get_state_full(swe_b200_handle* h, int buffer, void* data);
/* current_time / current_timestep (simulator.jl:141-142); stored in realtype like the reference */
int swe_b200_set_time(swe_b200_handle* h, double t, double dt);
int swe_b200_get_time(swe_b200_handle* h, double* t, double* dt);

/* perform_step!(sim, max_dt) (simulator.jl:145-177).  advance_time != 0 also does `t += dt`
 * (the loop body of simulate_to_time, :204-205).  dt stays on the device: no host sync. */
int swe_b200_step(swe_b200_handle* h, double max_dt, int advance_time);
/* n x { perform_step!(sim, min(maximum_timestep, +inf)); t += dt } back to back, no host sync */
int swe_b200_run_steps(swe_b200_handle* h, int64_t n, double maximum_timestep);
/* simulate_to_time(sim, t_end; match_endtime, maximum_timestep) (simulator.jl:179-212);
 * maximum_timestep <= 0 means `nothing`.  Blocks; returns the number of steps taken. */
int swe_b200_simulate_to_time(swe_b200_handle* h, double t_end, int match_endtime,
                              double maximum_timestep, int64_t* nsteps);

/* ---- fine-grained operators: the reference's per-function dispatch points ------------------ */
/* update_bc!(backend, grid, eq, data) (bc.jl:29-40, 49-60, 81-110) */
int swe_b200_update_bc(swe_b200_handle* h, int buffer);
/* reconstruct!(backend, rec, left, right, state, grid, eq, dir) (linearreconstruction.jl:58-102,
 * noreconstruction.jl:30-46); dir 0 = XDIR, 1 = YDIR.  Results via get_faces. */
int swe_b200_reconstruct(swe_b200_handle* h, int dir, int in_buffer);
int swe_b200_get_faces(swe_b200_handle* h, void* left, void* right);
/* add_time_derivative!(out, system, state, t) (system.jl:62-75): out is NOT zeroed first.
 * maxspeed[2] = maximum(wavespeeds) per direction (numericalflux.jl:32). */
int swe_b200_add_time_derivative(swe_b200_handle* h, int in_buffer, int out_buffer, double t,
                                 double maxspeed[2]);
int swe_b200_get_wavespeeds(swe_b200_handle* h, void* ws);   /* system.wavespeeds, (nx+2gc, ny+2gc) */
/* implicit_substep!(out, prev, system, dt) (friction.jl:80-84) */
int swe_b200_implicit_substep(swe_b200_handle* h, int out_buffer, int prev_buffer, double dt);
/* post_proc_substep!(out, system, eq) (timestepper/timestepper.jl:41-51) */
int swe_b200_post_proc_substep(swe_b200_handle* h, int buffer);

/* ---- the backend kernel surface's iteration spaces (backends/kernel_abstractions.jl:66-131,
 * grid.jl:213-253).  The precompiled backend cannot run arbitrary Julia closures; these entry
 * points run a device kernel over the same index space and record the 1-based indices each
 * work-item would receive, so the reference's test_for_each_*_2d.jl carry over.
 * out == NULL returns the count only. */
int swe_b200_enumerate_cells(swe_b200_handle* h, int64_t* out /*[n][2]*/, int64_t* count);
int swe_b200_enumerate_inner_cells(swe_b200_handle* h, int dir, int ghostcells,
                                   int64_t* out /*[n][6] left,mid,right*/, int64_t* count);
int swe_b200_enumerate_ghost_cells(swe_b200_handle* h, int dir, int64_t* out /*[n][2]*/,
                                   int64_t* count);

/* The same index spaces for a bare grid -- no simulator handle: this is what the backend-level kernel surface
 * `for_each_cell(f, backend, grid, y...)`, `for_each_inner_cell(f, backend, grid, direction, y...; ghostcells)`,
 * `for_each_ghost_cell(f, backend, grid, direction, y...)` (backends/kernel_abstractions.jl:73, 83, 129) binds.
 * kind 0 = cells, 1 = inner cells, 2 = ghost cells; nx, ny = interior cells, gc = grid.ghostcells; dir 0/1;
 * `ghostcells` is the keyword of for_each_inner_cell (ignored otherwise).  Errors: swe_b200_last_error(NULL). */
int swe_b200_enumerate_grid(int kind, int nx, int ny, int gc, int dir, int ghostcells, int device,
                            int64_t* out, int64_t* count);

/* ---- device-side reductions for callbacks ---------------------------------------------------------
 * Callbacks that only need scalars (examples/example_tools.jl:86-90 TotalWaterVolume:
 * sum(collect(current_interior_state(sim).h) .- B_cell); driven by IntervalWriter, callbacks.jl:40-55)
 * collect() the whole state in the reference.  This entry reduces over the interior cells of a buffer
 * on the device and returns 64 bytes:
 *   out[0] = sum(w - B_cell)   out[1] = sum(hu)   out[2] = sum(hv)   out[3] = max(w - B_cell)
 *   out[4] = min(w - B_cell)   out[5] = number of wet cells (w - B_cell >= depth_cutoff)
 *   out[6] = max(|hu|, |hv|)   out[7] = sum(w)
 * (depths in the storage type like the reference's broadcast, sums accumulated in Float64 in a fixed
 * order; of THIS handle's strip in a multi-GPU run). */
int swe_b200_diagnostics(swe_b200_handle* h, int buffer, double out[8]);

/* ---- asynchronous snapshots for writer callbacks -----------------------------------------------------
 * IntervalWriter-style callbacks (callbacks.jl:40-55; examples/example_tools.jl:113-137 writes w/hu/hv of the
 * interior every interval) block the reference on collect(state).  snapshot_begin copies the current state to a
 * device-side snapshot buffer on the compute stream (HBM-speed) and lets a copy stream move it to the host
 * arrays (nx x ny per variable, like get_state_interior) while the simulation continues; snapshot_wait blocks
 * until the host arrays are complete.  One snapshot in flight per handle (a second begin waits for the first).
 * For the copy to overlap, the host arrays should be page-locked: swe_b200_host_alloc/free wrap
 * cudaHostAlloc/cudaFreeHost for hosts that have no CUDA binding of their own. */
int swe_b200_snapshot_begin(swe_b200_handle* h, void* w, void* hu, void* hv);
int swe_b200_snapshot_wait(swe_b200_handle* h);
void* swe_b200_host_alloc(int64_t bytes);
int swe_b200_host_free(void* p);

/* ---- multi-GPU (y-strips; one process per GPU) -------------------------------------------- */
/* 128-byte descriptor (CUDA IPC handle + strip layout) of this rank's exchange window: the three
 * state volumes and the mailbox for halo stamps / CFL maxima.  Ranks exchange the descriptors with any
 * host-side all-gather (torch.distributed in the Python host, MPI in a Julia host). */
#define SWE_B200_IPC_BYTES 128
int swe_b200_get_ipc_handle(swe_b200_handle* h, void* blob128);
/* Map every rank's window (all_blobs: nranks x 128 B, ordered by rank).  Afterwards set_state_interior,
 * step, run_steps and simulate_to_time are COLLECTIVE: every rank must issue the same sequence. */
int swe_b200_open_peers(swe_b200_handle* h, const void* all_blobs128);
/* Unmap the peers' windows again (waits for this handle's stream first).  COLLECTIVE tear-down order: every rank calls
 * close_peers, then a host-side barrier, then destroy -- freeing a window that a peer still has mapped is undefined
 * behaviour in CUDA IPC.  Afterwards the handle is a single-rank handle without neighbours: only get_* / destroy make
 * sense on it. */
int swe_b200_close_peers(swe_b200_handle* h);

/* y-strips only.  Under PeriodicBC the first / last strip's low / high ghost rows -- and under NeumannBC the first strip's low
 * ghost rows (bc.jl:56) -- are an image of cell rows at the OTHER end of the domain and arrive by exchange; stage 2 rebuilds
 * U^1 in them and needs the bathymetry of those source cells.  `rows` = their (gc + 1) x (nx + 2gc + 1) corner values
 * (side 0: the corner rows of the top gc interior cell rows of the global domain, side 1: of the bottom gc interior rows),
 * element type cfg.dtype.  Required before the first step of such a strip when cfg.has_bathymetry. */
int swe_b200_set_wrap_bathymetry(swe_b200_handle* h, int side, const void* rows, int64_t n_elems);

/* ---- measurement ----------------------------------------------------------------------------- */
/* Runs n steps like run_steps and returns the device time (CUDA events on the handle's stream)
 * of the whole region and of the dominant stage kernels, plus the number of kernels launched. */
typedef struct swe_b200_timing {
    double total_ms;          /* events around the n steps                               */
    double stage_kernel_ms;   /* sum over the fused stage kernels (K1+K2) or 0 if unfused */
    int64_t launches;         /* kernels launched by the library in the region            */
    int64_t stage_launches;   /* number of stage-kernel launches timed                    */
    double stage1_ms;         /* sum over the stage-1 kernel launches (K1: U^n -> dU, CFL max) */
    double stage2_ms;         /* sum over the stage-2 kernel launches (K2: U^n, dU -> U^{n+1}) */
} swe_b200_timing;
int swe_b200_run_steps_timed(swe_b200_handle* h, int64_t n, double maximum_timestep,
                             swe_b200_timing* out);
/* total number of kernels this handle has launched so far */
int swe_b200_launch_count(swe_b200_handle* h, int64_t* count);

/* Which kernel family serves step / run_steps / simulate_to_time on this handle (>= 0; < 0 = error).  The
 * reference has one code path per operator (system.jl:62-75); this backend has several, selected at create()
 * from the configuration -- a host can log it, a test can assert it:
 *   PER_OPERATOR  one kernel per reference operator (ShallowWaterEquationsPure, gc != 2, cfg.fused == 0)
 *   FUSED_GENERIC the one-column fused stage kernels (strict policies; anything the packed kernels refuse)
 *   FUSED_X2      the packed two-column Float32 stage kernels (the performance path of the bench line)
 *   FUSED_D1      the Float64 ring / persistent-warp stage kernels */
enum { SWE_B200_FAMILY_PER_OPERATOR = 0, SWE_B200_FAMILY_FUSED_GENERIC = 1, SWE_B200_FAMILY_FUSED_X2 = 2,
       SWE_B200_FAMILY_FUSED_D1 = 3 };
int swe_b200_kernel_family(swe_b200_handle* h);

/* Per-cell comparison, ON THE DEVICE, of the states (buffer `buffer`) of two handles with the same grid, dtype
 * and device -- e.g. the fast-arithmetic handle against the strict (bit-faithful) one at sizes where the CPU
 * oracle cannot run and a D2H copy of both states would dominate.  Interior cells only (SURVEY 8c protocol):
 *   out[0..2] = max |a - b| of w, hu, hv      out[3..5] = max |b| of w, hu, hv   (b = `other`)
 *   out[6] = number of cells whose wet flag (w - B_cell >= depth_cutoff) differs
 *   out[7], out[8] = wet cells of a, of b      out[9] = cells holding a NaN in either state
 *   out[10..12] = number of cells whose |a - b| of w, hu, hv exceeds thresholds[0..2] (absolute values; a second
 *                 call with thresholds = tolerance x out[3..5] of the first counts the cells over a relative
 *                 tolerance); 0 when thresholds == NULL */
#define SWE_B200_COMPARE_SLOTS 13
int swe_b200_compare_states(swe_b200_handle* h, swe_b200_handle* other, int buffer, const double* thresholds,
                            double out[SWE_B200_COMPARE_SLOTS]);

#ifdef __cplusplus
}
#endif
#endif /* SWE_B200_H */

// engine.h -- host-side abstract interface between the C ABI (capi.cu) and the templated CUDA
// engine (engine_impl.cuh), which is instantiated per arithmetic policy in engine_fast.cu
// (native fp32/fp64 with FMA, -use_fast_math) and engine_strict.cu (-fmad=false, bit-faithful).
#pragma once
#include <cstdint>
#include <string>

#include "../../include/swe_b200.h"

namespace swe {

struct Engine {
    std::string err;
    virtual ~Engine() {}
    virtual int sync() = 0;
    virtual int set_bathymetry(const void* B, int64_t n) = 0;
    virtual int set_infiltration_factor(const double* f, int64_t n) = 0;
    virtual int set_rain_plane(const double* r, int64_t n) = 0;
    virtual int set_wrap_bathymetry(int side, const void* rows, int64_t n) = 0;
    virtual int set_state_interior(const void* w, const void* hu, const void* hv) = 0;
    virtual int get_state_interior(void* w, void* hu, void* hv) = 0;
    virtual int set_state_full(int buffer, const void* data) = 0;
    virtual int get_state_full(int buffer, void* data) = 0;
    virtual int set_time(double t, double dt) = 0;
    virtual int get_time(double* t, double* dt) = 0;
    virtual int step(double max_dt, int advance_time) = 0;
    virtual int run_steps(int64_t n, double maximum_timestep, swe_b200_timing* timing) = 0;
    virtual int simulate_to_time(double t_end, int match_endtime, double maximum_timestep, int64_t* nsteps) = 0;
    virtual int update_bc(int buffer) = 0;
    virtual int reconstruct(int dir, int in_buffer) = 0;
    virtual int get_faces(void* left, void* right) = 0;
    virtual int add_time_derivative(int in_buffer, int out_buffer, double t, double maxspeed[2]) = 0;
    virtual int get_wavespeeds(void* ws) = 0;
    virtual int diagnostics(int buffer, double out[8]) = 0;
    virtual int snapshot_begin(void* w, void* hu, void* hv) = 0;
    virtual int snapshot_wait() = 0;
    virtual int implicit_substep(int out_buffer, int prev_buffer, double dt) = 0;
    virtual int post_proc_substep(int buffer) = 0;
    virtual int enumerate(int kind, int dir, int ghostcells, int64_t* out, int64_t* count) = 0;
    virtual int get_ipc_handle(void* handle64) = 0;
    virtual int open_peers(const void* all_handles64) = 0;
    virtual int close_peers() = 0;
    virtual int64_t launch_count() const = 0;
    virtual int kernel_family() const = 0;
    virtual const void* state_ptr(int buffer) const = 0;
    virtual int elem_bytes() const = 0;
    virtual void dims(int out[4]) const = 0;          // nx, ny, gc, dtype
    virtual int compare(const void* other_state, int buffer, const double* thresholds, double out[SWE_B200_COMPARE_SLOTS]) = 0;
    virtual int device() const = 0;
};

// iteration spaces of the kernel surface for a bare grid (no simulator): swe_b200_enumerate_grid
int enumerate_grid(int kind, int nx, int ny, int gc, int dir, int ghostcells, int device, int64_t* out, int64_t* count,
                   std::string& err);

// factories, one per translation unit / arithmetic policy family
Engine* make_engine_fast(const swe_b200_config& cfg, std::string& err);
Engine* make_engine_strict(const swe_b200_config& cfg, std::string& err);

}  // namespace swe

// capi.cu -- the extern "C" boundary declared in include/swe_b200.h.
#include <cstring>
#include <string>

#include <cuda_runtime.h>

#include "engine.h"

struct swe_b200_handle {
    swe::Engine* e;
    std::string last;
};

static thread_local std::string g_create_error;

static int fail(swe_b200_handle* h, int rc) {
    if (h && h->e) h->last = h->e->err;
    return rc;
}
#define H_OR_FAIL(h) if (!(h) || !(h)->e) return SWE_B200_EINVAL

extern "C" {

int swe_b200_create(const swe_b200_config* c, swe_b200_handle** out) {
    if (!c || !out) return SWE_B200_EINVAL;
    *out = nullptr;
    auto bad = [&](const char* m, int rc) { g_create_error = m; return rc; };
    if (c->abi_version != SWE_B200_ABI_VERSION) return bad("abi_version mismatch", SWE_B200_EINVAL);
    if (c->dtype != SWE_B200_F32 && c->dtype != SWE_B200_F64) return bad("dtype must be F32 or F64", SWE_B200_EINVAL);
    if (c->nx < 1 || c->ny < 1 || c->gc < 1) return bad("nx, ny, gc must be positive", SWE_B200_EINVAL);
    if (c->bc < 0 || c->bc > 2) return bad("unknown boundary condition", SWE_B200_EINVAL);
    if (c->recon == SWE_B200_RECON_LINEAR && c->gc < 2)   // linearreconstruction.jl:45,59  @assert gc > 1
        return bad("LinearReconstruction needs ghostcells > 1", SWE_B200_EINVAL);
    if (c->eq_kind == 0 && c->has_bathymetry && !c->has_bottom_source)
        // equation.jl:74-84: non-zero bottom topography requires SourceTermBottom.  An all-zero array
        // would pass in the reference (is_zero); the host wrapper handles that case before create.
        return bad("Found non-zero bottom topography in equation, but no corresponding source term. "
                   "Did you forget to add a source term to your system?", SWE_B200_EINVAL);
    if (c->eq_kind == 0 && !c->has_bathymetry && c->const_B != 0.0 && !c->has_bottom_source)
        return bad("Found non-zero bottom topography in equation, but no corresponding source term. "
                   "Did you forget to add a source term to your system?", SWE_B200_EINVAL);
    if (c->eq_kind == 1 && (c->has_bathymetry || c->rain_kind || c->infil_kind))
        return bad("ShallowWaterEquationsPure carries no bathymetry / rain / infiltration on this path", SWE_B200_EUNSUPPORTED);
    if (c->n_rain < 0 || c->n_rain > 16) return bad("TimeDependentRain table too long (max 16)", SWE_B200_EUNSUPPORTED);
    if (c->nranks > 1 && (c->rank < 0 || c->rank >= c->nranks)) return bad("rank out of range", SWE_B200_EINVAL);
    std::string err;
    swe::Engine* e = c->strict ? swe::make_engine_strict(*c, err) : swe::make_engine_fast(*c, err);
    if (!e) { g_create_error = err; return SWE_B200_ECUDA; }
    *out = new swe_b200_handle{e, ""};
    return SWE_B200_OK;
}

int swe_b200_destroy(swe_b200_handle* h) {
    if (!h) return SWE_B200_EINVAL;
    if (h->e) cudaSetDevice(h->e->device());
    delete h->e;
    delete h;
    return SWE_B200_OK;
}
const char* swe_b200_last_error(swe_b200_handle* h) {
    if (!h) return g_create_error.c_str();
    if (h->e && !h->e->err.empty()) h->last = h->e->err;
    return h->last.c_str();
}
// every entry point runs with the handle's device current (several handles / devices per process are fine)
#define FWD(call) do { H_OR_FAIL(h); cudaSetDevice(h->e->device()); int rc_ = h->e->call; return rc_ ? fail(h, rc_) : rc_; } while (0)

int swe_b200_sync(swe_b200_handle* h) { FWD(sync()); }
int swe_b200_set_bathymetry(swe_b200_handle* h, const void* B, int64_t n) { FWD(set_bathymetry(B, n)); }
int swe_b200_set_infiltration_factor(swe_b200_handle* h, const double* f, int64_t n) { FWD(set_infiltration_factor(f, n)); }
int swe_b200_set_wrap_bathymetry(swe_b200_handle* h, int side, const void* rows, int64_t n) { FWD(set_wrap_bathymetry(side, rows, n)); }
int swe_b200_set_rain_plane(swe_b200_handle* h, const double* r, int64_t n) { FWD(set_rain_plane(r, n)); }
int swe_b200_set_state_interior(swe_b200_handle* h, const void* w, const void* hu, const void* hv) { FWD(set_state_interior(w, hu, hv)); }
int swe_b200_get_state_interior(swe_b200_handle* h, void* w, void* hu, void* hv) { FWD(get_state_interior(w, hu, hv)); }
int swe_b200_set_state_full(swe_b200_handle* h, int b, const void* d) { FWD(set_state_full(b, d)); }
int swe_b200_get_state_full(swe_b200_handle* h, int b, void* d) { FWD(get_state_full(b, d)); }
int swe_b200_set_time(swe_b200_handle* h, double t, double dt) { FWD(set_time(t, dt)); }
int swe_b200_get_time(swe_b200_handle* h, double* t, double* dt) { FWD(get_time(t, dt)); }
int swe_b200_step(swe_b200_handle* h, double max_dt, int adv) { FWD(step(max_dt, adv)); }
int swe_b200_run_steps(swe_b200_handle* h, int64_t n, double mts) { FWD(run_steps(n, mts, nullptr)); }
int swe_b200_run_steps_timed(swe_b200_handle* h, int64_t n, double mts, swe_b200_timing* t) { FWD(run_steps(n, mts, t)); }
int swe_b200_simulate_to_time(swe_b200_handle* h, double t_end, int match, double mts, int64_t* ns) { FWD(simulate_to_time(t_end, match, mts, ns)); }
int swe_b200_update_bc(swe_b200_handle* h, int b) { FWD(update_bc(b)); }
int swe_b200_reconstruct(swe_b200_handle* h, int dir, int b) { FWD(reconstruct(dir, b)); }
int swe_b200_get_faces(swe_b200_handle* h, void* l, void* r) { FWD(get_faces(l, r)); }
int swe_b200_add_time_derivative(swe_b200_handle* h, int ib, int ob, double t, double ms[2]) { FWD(add_time_derivative(ib, ob, t, ms)); }
int swe_b200_get_wavespeeds(swe_b200_handle* h, void* w) { FWD(get_wavespeeds(w)); }
int swe_b200_diagnostics(swe_b200_handle* h, int b, double out[8]) { FWD(diagnostics(b, out)); }
int swe_b200_snapshot_begin(swe_b200_handle* h, void* w, void* hu, void* hv) { FWD(snapshot_begin(w, hu, hv)); }
int swe_b200_snapshot_wait(swe_b200_handle* h) { FWD(snapshot_wait()); }
void* swe_b200_host_alloc(int64_t bytes) {
    void* p = nullptr;
    if (bytes <= 0 || cudaHostAlloc(&p, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
int swe_b200_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? SWE_B200_OK : SWE_B200_ECUDA; }
int swe_b200_implicit_substep(swe_b200_handle* h, int ob, int pb, double dt) { FWD(implicit_substep(ob, pb, dt)); }
int swe_b200_post_proc_substep(swe_b200_handle* h, int b) { FWD(post_proc_substep(b)); }
int swe_b200_enumerate_cells(swe_b200_handle* h, int64_t* out, int64_t* n) { FWD(enumerate(0, 0, 0, out, n)); }
int swe_b200_enumerate_inner_cells(swe_b200_handle* h, int dir, int g, int64_t* out, int64_t* n) { FWD(enumerate(1, dir, g, out, n)); }
int swe_b200_enumerate_ghost_cells(swe_b200_handle* h, int dir, int64_t* out, int64_t* n) { FWD(enumerate(2, dir, 0, out, n)); }
int swe_b200_enumerate_grid(int kind, int nx, int ny, int gc, int dir, int ghostcells, int device, int64_t* out, int64_t* n) {
    if (!n || kind < 0 || kind > 2 || dir < 0 || dir > 1 || nx < 1 || ny < 1 || gc < 0 || ghostcells < 0) return SWE_B200_EINVAL;
    std::string err;
    int rc = swe::enumerate_grid(kind, nx, ny, gc, dir, ghostcells, device, out, n, err);
    if (rc) g_create_error = err;
    return rc;
}
int swe_b200_get_ipc_handle(swe_b200_handle* h, void* p) { FWD(get_ipc_handle(p)); }
int swe_b200_open_peers(swe_b200_handle* h, const void* p) { FWD(open_peers(p)); }
int swe_b200_close_peers(swe_b200_handle* h) { FWD(close_peers()); }
int swe_b200_kernel_family(swe_b200_handle* h) { H_OR_FAIL(h); return h->e->kernel_family(); }
int swe_b200_compare_states(swe_b200_handle* h, swe_b200_handle* other, int buffer, const double* thresholds,
                            double out[SWE_B200_COMPARE_SLOTS]) {
    H_OR_FAIL(h);
    H_OR_FAIL(other);
    int a[4], b[4];
    h->e->dims(a);
    other->e->dims(b);
    if (a[0] != b[0] || a[1] != b[1] || a[2] != b[2] || a[3] != b[3] || h->e->device() != other->e->device()) {
        h->last = h->e->err = "compare_states: the two handles must have the same grid, dtype and device";
        return SWE_B200_EINVAL;
    }
    if (buffer < 0 || buffer > 2) return SWE_B200_EINVAL;
    cudaSetDevice(h->e->device());
    int rc = other->e->sync();                        // the other handle works on its own stream
    if (rc) return fail(other, rc);
    rc = h->e->compare(other->e->state_ptr(buffer), buffer, thresholds, out);
    return rc ? fail(h, rc) : rc;
}
int swe_b200_launch_count(swe_b200_handle* h, int64_t* c) { H_OR_FAIL(h); *c = h->e->launch_count(); return SWE_B200_OK; }

}  // extern "C"

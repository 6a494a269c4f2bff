// engine_fast.cu -- native-arithmetic instantiations (FMA contraction on, -use_fast_math for fp32).
#define SWE_FAST_MINMAX 1      // Float64 min/max as compare-and-select (physics.cuh)
#define SWE_FAST_DIV64 1       // Float64 reciprocals: hardware seed + two Newton steps (physics.cuh)
#define SWE_TU_NS tu_fast          // tu_namespace.h: per-TU inline namespace (no ODR clash between the policies)
#include "engine_impl.cuh"
namespace swe {
Engine* make_engine_fast(const swe_b200_config& cfg, std::string& err) {
    if (cfg.dtype == SWE_B200_F32) return make_engine_t<Policy<float, float, false>>(cfg, err);
    return make_engine_t<Policy<double, double, false>>(cfg, err);
}
// kernel_abstractions.jl:66-131 index spaces of a grid that has no simulator yet (the Julia-side for_each_* methods
// of B200Backend take (backend, grid), not a simulator): same kernel as EngineT::enumerate.
int enumerate_grid(int kind, int nx, int ny, int gc, int dir, int g, int device, int64_t* out, int64_t* count, std::string& err) {
    const int Nx = nx + 2 * gc, Ny = ny + 2 * gc;
    long long na, nb;
    int per;
    if (kind == 0) { na = Nx; nb = Ny; per = 2; }
    else if (kind == 1) { na = dir == 0 ? Nx - 2 * g : Nx - 2 * gc; nb = dir == 0 ? Ny - 2 * gc : Ny - 2 * g; per = 6; }
    else { na = dir == 0 ? gc : Nx - 2 * gc; nb = dir == 0 ? Ny - 2 * gc : gc; per = 2; }
    if (na < 0 || nb < 0) { err = "empty index space"; return SWE_B200_EINVAL; }
    *count = na * nb;
    if (!out || *count == 0) return SWE_B200_OK;
    auto fail = [&](const char* what, cudaError_t e) { err = std::string(what) + ": " + cudaGetErrorString(e); return SWE_B200_ECUDA; };
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return fail("cudaSetDevice", e);
    long long* d = nullptr;
    e = cudaMalloc(&d, sizeof(long long) * per * (*count));
    if (e != cudaSuccess) return fail("cudaMalloc", e);
    dim3 b(32, 8), gr((unsigned)((na + 31) / 32), (unsigned)((nb + 7) / 8));
    k_enumerate<<<gr, b>>>(kind, dir, g, Nx, Ny, gc, d);
    e = cudaMemcpy(out, d, sizeof(long long) * per * (*count), cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail("k_enumerate", e);
    return SWE_B200_OK;
}
}  // namespace swe

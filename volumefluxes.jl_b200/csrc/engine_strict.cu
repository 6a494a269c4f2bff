// engine_strict.cu -- bit-faithful instantiations; this file is compiled with -fmad=false and
// IEEE division / square root so that every operation is the literal sequence of the reference.
// f32: the reference's mixed Float32-storage / Float64-scalar promotion.
#define SWE_TU_NS tu_strict          // tu_namespace.h: per-TU inline namespace (no ODR clash between the policies)
#include "engine_impl.cuh"
namespace swe {
Engine* make_engine_strict(const swe_b200_config& cfg, std::string& err) {
    if (cfg.dtype == SWE_B200_F32) return make_engine_t<Policy<float, double, true>>(cfg, err);
    return make_engine_t<Policy<double, double, true>>(cfg, err);
}
}  // namespace swe

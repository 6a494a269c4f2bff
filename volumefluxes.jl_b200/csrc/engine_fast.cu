// engine_fast.cu -- native-arithmetic instantiations (FMA contraction on, -use_fast_math for fp32).
#define SWE_FAST_MINMAX 1      // Float64 min/max as compare-and-select (physics.cuh)
#define SWE_FAST_DIV64 1       // Float64 reciprocals: hardware seed + two Newton steps (physics.cuh)
#include "engine_impl.cuh"
namespace swe {
Engine* make_engine_fast(const swe_b200_config& cfg, std::string& err) {
    if (cfg.dtype == SWE_B200_F32) return make_engine_t<Policy<float, float, false>>(cfg, err);
    return make_engine_t<Policy<double, double, false>>(cfg, err);
}
}  // namespace swe

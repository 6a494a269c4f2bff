// tu_namespace.h -- every kernel / device function of the engine lives in an inline namespace named after the
// translation unit that compiles it.  engine_fast.cu (-use_fast_math, ftz) and engine_strict.cu (-fmad=false,
// IEEE div/sqrt) include the same templates; without this, templates that do not depend on the arithmetic
// policy (k_bc<float>, k_fill<float>, bc_source, ...) would have ONE mangled name compiled twice with different
// floating-point flags, and the linker would keep whichever copy it saw first (an ODR violation that could hand
// the strict engine a flush-to-zero kernel).
#pragma once
#ifndef SWE_TU_NS
#error "define SWE_TU_NS (a distinct name per translation unit) before including the engine headers"
#endif
#define SWE_NS_BEGIN namespace swe { inline namespace SWE_TU_NS {
#define SWE_NS_END } }

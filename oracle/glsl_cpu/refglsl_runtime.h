// oracle/glsl_cpu/refglsl_runtime.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Program registry of librefglsl.so: every reference shader program is one translation unit (refglsl_prog.cpp compiled
// with -DPROG_NAME / -DPROG_HEADER) that registers a dispatch function here.
#pragma once
#include "glsl_emu.h"

namespace refglsl {
typedef void (*DispatchFn)(glsl::UniformTable* uniforms, const int groups[3], const int local[3]);
void register_program(const char* name, DispatchFn fn);
}  // namespace refglsl

// oracle/ref_stubs/msvc_compat.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
// Force-included (-include) only for the reference's libs/file_utils/pvm.cpp, which is written against the MSVC
// "secure CRT" names.  Nothing of the reference is copied: these are three spellings mapped to their ISO C twins.
#pragma once
#include <cassert>
#include <cstdio>
typedef int errno_t;
static inline errno_t fopen_s(FILE** f, const char* name, const char* mode) { *f = std::fopen(name, mode); return *f ? 0 : 1; }
#define sscanf_s sscanf

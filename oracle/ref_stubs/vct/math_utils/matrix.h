// oracle/ref_stubs/vct -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Empty stand-in for the reference's
// libs/math_utils/matrix.h (MSVC-only templates), on the include path only while rc1pvctsg/preprocessingstages.cpp is
// compiled in place for oracle/_ref/libref.so; nothing in that translation unit uses it.
#pragma once

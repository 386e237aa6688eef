// oracle/ref_stubs/vct -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Stand-in for libs/math_utils/matrix4.h: gl_utils/shader.h
// only names the type in one declaration.
#pragma once
namespace lqc { class Matrix4f; }
using lqc::Matrix4f;

// oracle/ref_stubs/math_utils/utils.h -- TEST INFRASTRUCTURE.
// Stand-in for the reference's libs/math_utils/utils.h, which drags in MSVC-only matrix templates.
// It is put FIRST on the include path when oracle/_ref/libref.so compiles the reference's conegaussiansampler.cpp in
// place; it only declares what that file uses.  RodriguesRotation is defined in oracle/ref_shim.cpp as a restatement
// of libs/math_utils/utils.cpp:149-165.
#ifndef VRB_REF_STUB_MATH_UTILS_H
#define VRB_REF_STUB_MATH_UTILS_H
#include <cmath>
#include <glm/glm.hpp>
#include <glm/gtc/constants.hpp>
#ifndef DEGREE_TO_RADIANS
#define DEGREE_TO_RADIANS(s) (s * (glm::pi<double>() / 180.0))
#endif
glm::vec3 RodriguesRotation(glm::vec3 v, float teta, glm::vec3 k);
glm::dvec3 RodriguesRotation(glm::dvec3 v, double teta, glm::dvec3 k);
#endif

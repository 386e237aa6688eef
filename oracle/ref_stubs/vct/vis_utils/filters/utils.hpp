// oracle/ref_stubs/vct -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Empty stand-in for libs/vis_utils/filters/utils.hpp
// (does not compile with g++); pulled in by renderoutputframe.h, unused by preprocessingstages.cpp.
#pragma once

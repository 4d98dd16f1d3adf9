// Force-included before the reference sources: MSVC accepts std::powf / std::fmaxf, libstdc++ does not
// (reference uses them at triangle_math.cpp:55 and Lightning.h:177).  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <cmath>
#include <cstring>
#include <cstdio>
#include <algorithm>
#include <string>
namespace std { using ::powf; using ::fmaxf; }

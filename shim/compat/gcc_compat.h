// Force-included when the reference's MeshS / Material / rgb2spec sources are compiled with g++ for the shim demo:
// MSVC accepts std::powf / std::fmaxf, libstdc++ does not (the reference uses them at triangle_math.cpp:55 and
// Lightning.h:177).  A maintainer building with MSVC does not need this header.
#pragma once
#include <cmath>
#include <cstring>
#include <cstdio>
#include <algorithm>
#include <string>
namespace std { using ::powf; using ::fmaxf; }

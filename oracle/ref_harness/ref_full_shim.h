// TEST INFRASTRUCTURE ONLY (oracle/Makefile.full, passed with -include). nvcc 12.9's front end resolves the unqualified call to
// parallel_for_gpu in include/neural-graphics-primitives/takikawa_encoding.cuh:389 at template definition time, where the reference
// relies on the name being visible only later (`using namespace tcnn` in the including .cu). Making the name visible in ngp:: up front
// compiles the unmodified reference source.
#pragma once
#include <tiny-cuda-nn/common.h>
namespace ngp { using tcnn::parallel_for_gpu; }

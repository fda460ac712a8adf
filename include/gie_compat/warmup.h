#pragma once
#include "cuda_toolkit/cuda_macro.h"
inline void warmupCuda() { GIE_CHECK(gie_warmup()); }   // include/warmup.h:9, src/kernel/edt/warmup.cu

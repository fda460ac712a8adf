#pragma once
// GIE_COMPAT_REFERENCE_MAPMAKERS: keep the reference's own src/*_map_maker.cpp (declarations only here); default: the
// header-only MapMakers of map_makers.h, which need no .cpp and no device staging buffers.
#ifdef GIE_COMPAT_REFERENCE_MAPMAKERS
#include "cuda_toolkit/occupancy/map_makers_decl.h"
#else
#include "cuda_toolkit/occupancy/map_makers.h"
#endif

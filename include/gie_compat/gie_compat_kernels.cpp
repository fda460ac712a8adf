// The one translation unit a build adds when it keeps the reference's own src/*_map_maker.cpp: it emits
// PNTCLD_RAYCAST / HOKUYO_FAST / VLP_FAST / REALSENSE_FAST ::localOGMKernels as linkable symbols (bodies in
// kernel/ogm_interfaces.h: one C-ABI call each), replacing src/kernel/{point_cloud,hokuyo,vlp16,realsense}/*.cu.
#define GIE_COMPAT_EMIT_KERNEL_SYMBOLS
#include "kernel/ogm_interfaces.h"

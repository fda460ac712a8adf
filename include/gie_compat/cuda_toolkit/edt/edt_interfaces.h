// Stand-in for include/cuda_toolkit/edt/edt_interfaces.h.
#pragma once
#include "cutt/cutt.h"
#include "map_structure/local_batch.h"
namespace EDT_OCC {
// src/kernel/edt/local_edt.cu:7-28.  `plan` (three cuTT handles) and `time` are accepted and unused.
inline void batchEDTUpdate(LocMap *loc_map, cuttHandle * /*plan*/, const int /*time*/)
{
    GIE_CHECK(gie_edt_batch_update(loc_map->handle()));
}
}  // namespace EDT_OCC

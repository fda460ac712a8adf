#pragma once
#include "cuda_toolkit/occupancy/map_makers.h"

#pragma once
#include <vector_types.h>
#include "cuda_toolkit/occupancy/sensor_params.h"

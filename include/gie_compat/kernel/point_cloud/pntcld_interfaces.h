#pragma once
#include "kernel/ogm_interfaces.h"

// pack_kernel.cuh — LZ4 frame encoder kernel (K2).  Filled in by the pack milestone.
#pragma once
#include "common.cuh"

// Drop-in replacement for GS-LIVM include/gs/cuda_rasterizer/config.h:15-17 (same three macros; the
// reference's operator shim reads NUM_CHANNELS at src/gs/rasterize_points.cu:78,175). Values are fixed
// by the kernels in libgsr_b200.so and mirrored by GSR_* in ../gsr.h.
#pragma once
#include "../gsr.h"

#define NUM_CHANNELS GSR_NUM_CHANNELS
#define BLOCK_X GSR_BLOCK_X
#define BLOCK_Y GSR_BLOCK_Y

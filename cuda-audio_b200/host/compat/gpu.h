/* source-compatibility shim for src/gpu.h: selectGpu() (gpu.cu:38-90) asserts on sm_100; here it just binds device 0 */
#pragma once
#include <cuda_runtime.h>
#include <cassert>
#include <cstdint>
#include <string>
#include "../logger.h"
inline void selectGpu() { cudaSetDevice(0); }

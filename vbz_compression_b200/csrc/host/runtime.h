// csrc/host/runtime.h - CUDA runtime access for the host layer.
#pragma once
#ifdef VBZ_CPUSIM
#include "cuda_runtime_sim.h"  // tests/cpusim: emulator build used by the CPU test-suite only
#else
#include <cuda_runtime.h>
#endif

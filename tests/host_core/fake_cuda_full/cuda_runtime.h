/* stands in for <cuda_runtime.h> when a .cu file is compiled for the host under tests/host_core/cuda_emu_full.h (see that file) */
#include "../cuda_emu_full.h"

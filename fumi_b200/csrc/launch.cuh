// Kernel-launch indirection.  In the product build this is a plain <<<>>> launch.  The test-only
// host emulation (tests/emu/, -DFUMI_EMU) runs the same kernel bodies on CPU threads so their
// indexing / synchronisation logic can be exercised where no GPU exists; it is never shipped.
#pragma once

#ifdef FUMI_EMU
#include "cuda_emu.h"
#define FUMI_LAUNCH(kernel, grid, block, smem, stream, ...) \
    fumi_emu::launch([&]() { kernel(__VA_ARGS__); }, (grid), (block), (smem))
#else
#include <cuda_runtime.h>
#define FUMI_DYN_SMEM(type, name) extern __shared__ __align__(16) type name[]
#define FUMI_LAUNCH(kernel, grid, block, smem, stream, ...) \
    kernel<<<(grid), (block), (smem), (cudaStream_t)(stream)>>>(__VA_ARGS__)
#endif

// TEST INFRASTRUCTURE (see cuda_emu.h): storage of the CUDA built-in variables.  Kernel code sees them as
// `extern const` (each fiber observes constant values); only the scheduler in cuda_emu.cpp writes them.
#include <vector_types.h>

uint3 threadIdx, blockIdx;
dim3 blockDim, gridDim;

void gv_emu_set_thread_index(unsigned x, unsigned y, unsigned z) {
    threadIdx.x = x, threadIdx.y = y, threadIdx.z = z;
}
void gv_emu_set_block_index(unsigned x, unsigned y, unsigned z) {
    blockIdx.x = x, blockIdx.y = y, blockIdx.z = z;
}
void gv_emu_set_dimensions(dim3 grid, dim3 block) {
    gridDim = grid;
    blockDim = block;
}

// TEST INFRASTRUCTURE (see oracle/emulate_reference.py): what the reference expects from nvcc's environment
// beyond tests/emu/cuda_emu.h, force-included in front of its mirrored sources.
#pragma once
#include <curand.h>


// Under the CUDA emulation the reference's device generator is cuRAND's HOST generator -- the same XORWOW stream
// for the same seed (cuRAND's documented guarantee, checked on a B200: tests/golden/curand.npz) written into what
// the emulation calls device memory.
#define curandCreateGenerator curandCreateGeneratorHost
#define curandSetStream(generator, stream) CURAND_STATUS_SUCCESS

// CUDA declares float overloads of the C math functions in the global namespace and the reference's device code
// relies on them (abs(float) would otherwise be the integer abs, sqrt(float) a double computation): libstdc++'s
// <math.h> / <stdlib.h> wrappers import the std:: overload sets into the global namespace.
#include <math.h>
#include <stdlib.h>

// the device pass of the two-pass build (see oracle/Makefile): defined only after the CUDA headers were read as
// host code
#ifdef GV_EMU_DEVICE_PASS
#define __CUDA_ARCH__ 1000
#endif
#define __CUDACC_VER_MAJOR__ 12  // the reference picks the *_sync shuffles from it (util/gpu.cuh:51-65)

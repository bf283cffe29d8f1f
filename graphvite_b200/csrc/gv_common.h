// Shared helpers for the libgv_b200 translation units (host side; no CUDA headers needed).
#pragma once

#include <string>

#include "gv_b200.h"

namespace gv {

// thread-local error slot behind gv_last_error()
void set_error(const std::string &message);
int fail(const std::string &message);  // set_error + return -1
bool direct_fill_per_walk();           // tunable `fill_per_walk` (gv_train.cu): 1 = the thread-per-walk fill at P = 1
int sampler_max_ctas();                // tunable `sampler_max_ctas` (gv_train.cu): grid cap of the walk kernels, 0 = none
int kg_kernel_flags();                 // tunable `kg_flags` (gv_kg.cu): bit 0 = IEEE sqrt / division / sincosf in kg_train_kernel
void set_kg_kernel_flags(int value);

}  // namespace gv

#define GV_CUDA_OK(call)                                                                          \
    do {                                                                                          \
        cudaError_t gv_err__ = (call);                                                            \
        if (gv_err__ != cudaSuccess)                                                              \
            return gv::fail(std::string("CUDA error ") + cudaGetErrorString(gv_err__) + " at " +  \
                            __FILE__ + ":" + std::to_string(__LINE__));                           \
    } while (0)

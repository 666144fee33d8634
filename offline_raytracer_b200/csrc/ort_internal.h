// ort_internal.h -- what the translation units of libort_b200.so share besides the public
// headers: the per-thread error string and the CUDA error macro.
#pragma once

#include <string>
#include <cuda_runtime.h>

#include "ort_b200.h"

extern "C" void ort_set_last_error_(const char *msg);   // ort_b200.cu

namespace ort {

inline int fail_with(int code, const std::string &msg)
{
    ort_set_last_error_(msg.c_str());
    return code;
}

} // namespace ort

#define ORT_CUDA_TRY(expr)                                                                              \
    do {                                                                                                \
        cudaError_t e__ = (expr);                                                                       \
        if(e__ != cudaSuccess)                                                                          \
            return ort::fail_with(ORT_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__));   \
    } while(0)

// Every extern "C" body that can allocate runs under this guard: no C++ exception crosses the C ABI
// (the reference's own failure mode is an assert / null dereference, code/platform.h:16-20; this
// library reports instead)
#define ORT_GUARD_BEGIN try {
#define ORT_GUARD_END                                                                                   \
    } catch(const std::bad_alloc &) { return ort::fail_with(ORT_ERR_LIMIT, "out of host memory"); }    \
      catch(const std::exception &e__) { return ort::fail_with(ORT_ERR_ARG, std::string("exception: ") + e__.what()); } \
      catch(...) { return ort::fail_with(ORT_ERR_ARG, "unknown exception"); }

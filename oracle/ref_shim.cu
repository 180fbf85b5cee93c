// ref_shim.cu — TEST INFRASTRUCTURE.  extern "C" doors onto the UNMODIFIED reference kernels.
//
// Compiled together with /root/reference/kernel/csrc/SpMM_API.cu (sources stay where they are; only
// the built oracle/_ref/libmustafar_ref.so travels to the GPU box).  The reference API has C++
// linkage (kernel/build/SpMM_API.cuh:46-64, :92-110); these wrappers pass the arguments through in the
// same way kernel/kernel_wrapper/mustafar_wrapper.cu:113-131 and :242-260 do (A = NULL, N = 8,
// Split_K = 1) and return the cudaError_t the reference wrapper drops.
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "SpMM_API.cuh"

extern "C" int ref_key_formulation(void* stream, const uint64_t* bmp, const void* NZ, const uint32_t* idx,
                                   const uint32_t* NZ_offset, const void* B, void* C, int M_Global, int K_Global,
                                   int Batch_Size, int num_key_value_groups) {
    return static_cast<int>(Key_SplitK_API(static_cast<cudaStream_t>(stream), static_cast<const half*>(nullptr), bmp,
                                           static_cast<const uint4*>(NZ), idx, NZ_offset, static_cast<const half*>(B),
                                           static_cast<half*>(C), M_Global, 8, K_Global, static_cast<half*>(nullptr), 1,
                                           Batch_Size, num_key_value_groups));
}

extern "C" int ref_value_formulation(void* stream, const uint64_t* bmp, const void* NZ, const uint32_t* idx,
                                     const uint32_t* NZ_offset, const void* B, void* C, void* Reduction_Workspace,
                                     int M_Global, int K_Global, int Batch_Size, int num_key_value_groups) {
    return static_cast<int>(Value_SplitK_API(static_cast<cudaStream_t>(stream), static_cast<const half*>(nullptr), bmp,
                                             static_cast<const uint4*>(NZ), idx, NZ_offset, static_cast<const half*>(B),
                                             static_cast<half*>(C), M_Global, 8, K_Global,
                                             static_cast<half*>(Reduction_Workspace), 1, Batch_Size,
                                             num_key_value_groups));
}

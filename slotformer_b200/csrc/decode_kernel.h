// Internal interface between the C ABI (capi.cu) and the decoder-epilogue kernels.
#pragma once
#include <cuda_runtime.h>

namespace sfb {

// masks = softmax over slots of plane 3, recon = sum_k planes 0..2 * masks; slot_max (nullable) receives the
// bit patterns of max_pixels masks[b][k] (must be zeroed by the caller)
cudaError_t decode_combine_launch(const float* dec, float* masks, float* recon, unsigned int* slot_max, int B, int K, int HW,
                                  int sms, cudaStream_t st);
cudaError_t mask_max_launch(const float* masks, unsigned int* slot_max, int BK, int HW, int sms, cudaStream_t st);
cudaError_t seg_argmax_launch(const float* masks, const unsigned int* slot_max, long long* seg, int B, int K, int HW,
                              float fg_thre, int sms, cudaStream_t st);

}  // namespace sfb

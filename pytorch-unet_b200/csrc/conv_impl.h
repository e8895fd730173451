// Internal interface between the C-ABI dispatch (conv_api.cu) and the two implementations of the convolution
// family: CUDA-core "direct" kernels (conv_direct.cu) and tcgen05/TMEM/TMA kernels (conv_umma.cu, wgrad_umma.cu).
#pragma once
#include "common.cuh"

namespace b200 {

int direct_conv_fwd(const b200_conv_fwd_params* p, cudaStream_t st);
int direct_conv_dgrad(const b200_conv_dgrad_params* p, cudaStream_t st);
int direct_conv_wgrad(const b200_conv_wgrad_params* p, void* ws, size_t ws_bytes, cudaStream_t st);
int direct_convt_fwd(const b200_convt_fwd_params* p, cudaStream_t st);
int direct_convt_dgrad(const b200_convt_dgrad_params* p, cudaStream_t st);
int direct_convt_wgrad(const b200_convt_wgrad_params* p, void* ws, size_t ws_bytes, cudaStream_t st);
size_t direct_wgrad_workspace(long long npix, int cout, int cin_total, int taps);
int bias_grad(const b200_view& dz, float* db, void* ws, cudaStream_t st);

// first-layer kernels (1..4 input channels), chosen by AUTO when the tcgen05 path cannot take the shape
bool smallc_conv_fwd_ok(const b200_conv_fwd_params* p);
int smallc_conv_fwd(const b200_conv_fwd_params* p, cudaStream_t st);
bool smallc_conv_wgrad_ok(const b200_conv_wgrad_params* p);
size_t smallc_conv_wgrad_workspace(const b200_conv_wgrad_params* p);
int smallc_conv_wgrad(const b200_conv_wgrad_params* p, void* ws, size_t ws_bytes, cudaStream_t st);

// tcgen05 path.  *_ok() say whether the shape can be taken (channel multiples, alignment); the launchers assume it.
bool umma_conv_fwd_ok(const b200_conv_fwd_params* p);
bool umma_conv_dgrad_ok(const b200_conv_dgrad_params* p);
bool umma_conv_wgrad_ok(const b200_conv_wgrad_params* p);
bool umma_convt_fwd_ok(const b200_convt_fwd_params* p);
bool umma_convt_dgrad_ok(const b200_convt_dgrad_params* p);
bool umma_convt_wgrad_ok(const b200_convt_wgrad_params* p);
int umma_conv_fwd(const b200_conv_fwd_params* p, cudaStream_t st);
int umma_conv_dgrad(const b200_conv_dgrad_params* p, cudaStream_t st);
int umma_conv_wgrad(const b200_conv_wgrad_params* p, void* ws, size_t ws_bytes, cudaStream_t st);
size_t umma_conv_wgrad_workspace(const b200_conv_wgrad_params* p);
int umma_convt_fwd(const b200_convt_fwd_params* p, cudaStream_t st);
int umma_convt_dgrad(const b200_convt_dgrad_params* p, cudaStream_t st);
int umma_convt_wgrad(const b200_convt_wgrad_params* p, void* ws, size_t ws_bytes, cudaStream_t st);
size_t umma_convt_wgrad_workspace(const b200_convt_wgrad_params* p);

}  // namespace b200

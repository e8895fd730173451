// tcgen05 backward-weights (placeholder until the kernel lands): reports every shape as unsupported.
#include "conv_impl.h"
namespace b200 {
bool umma_conv_wgrad_ok(const b200_conv_wgrad_params*) { return false; }
bool umma_convt_wgrad_ok(const b200_convt_wgrad_params*) { return false; }
int umma_conv_wgrad(const b200_conv_wgrad_params*, void*, size_t, cudaStream_t) { return fail(-1, "not built"); }
size_t umma_conv_wgrad_workspace(const b200_conv_wgrad_params*) { return 0; }
int umma_convt_wgrad(const b200_convt_wgrad_params*, void*, size_t, cudaStream_t) { return fail(-1, "not built"); }
size_t umma_convt_wgrad_workspace(const b200_convt_wgrad_params*) { return 0; }
}  // namespace b200

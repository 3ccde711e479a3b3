// placeholder until the tcgen05 kernels land
#include "common.cuh"
namespace nb200
{
    bool tc_forward_supported(const nb200_conv_desc&) { return false; }
    bool tc_input_gradient_supported(const nb200_conv_desc&) { return false; }
    bool tc_kernels_gradient_supported(const nb200_conv_desc&) { return false; }
    size_t tc_workspace_bytes(int, const nb200_conv_desc&) { return 0; }
    int tc_forward(const nb200_conv_desc&, const float*, const float*, const float*, int, float, float*, void*, size_t, cudaStream_t) { return fail(NB200_E_UNSUPPORTED, "tc"); }
    int tc_input_gradient(const nb200_conv_desc&, const float*, const float*, float*, void*, size_t, cudaStream_t) { return fail(NB200_E_UNSUPPORTED, "tc"); }
    int tc_kernels_gradient(const nb200_conv_desc&, const float*, const float*, float*, void*, size_t, cudaStream_t) { return fail(NB200_E_UNSUPPORTED, "tc"); }
}

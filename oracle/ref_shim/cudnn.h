// Test infrastructure only. Minimal stand-in for <cudnn.h>: the reference's Shape.cpp
// (Shape.cpp:2,79,228-229) only creates/destroys a 4-d tensor descriptor, which the CPU
// convolution path never reads.
#pragma once
typedef struct cudnnTensorStruct* cudnnTensorDescriptor_t;
enum { CUDNN_TENSOR_NCHW = 0 };
enum { CUDNN_DATA_FLOAT = 0 };
static inline int cudnnCreateTensorDescriptor(cudnnTensorDescriptor_t* d) { *d = nullptr; return 0; }
static inline int cudnnDestroyTensorDescriptor(cudnnTensorDescriptor_t) { return 0; }
static inline int cudnnSetTensor4dDescriptor(cudnnTensorDescriptor_t, int, int, int, int, int, int) { return 0; }

// placeholder until the tcgen05 path lands
#pragma once
#include "common.cuh"
namespace clairb { namespace tc {
struct Weights {}; struct Workspace {};
struct HostModel { const float* lstm_kernel[2][2]; const float* lstm_bias[2][2]; const float *w3, *b3, *W4, *b4; float *d_W5, *d_b5, *d_Whd, *d_bhd; };
inline bool available() { return false; }
inline cudaError_t build_weights(Weights&, const HostModel&) { return cudaErrorNotSupported; }
inline void free_weights(Weights&) {}
inline cudaError_t alloc_workspace(Workspace&, int64_t) { return cudaErrorNotSupported; }
inline void free_workspace(Workspace&) {}
inline cudaError_t forward(const Weights&, Workspace&, const void*, int, SiteMap, float*, float*, cudaStream_t, int*) { return cudaErrorNotSupported; }
inline cudaError_t get_layer(Workspace&, int, SiteMap, float*) { return cudaErrorNotSupported; }
}}

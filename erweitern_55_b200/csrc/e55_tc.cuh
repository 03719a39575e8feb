// Tensor-core (tcgen05) path -- placeholder until the tower kernel lands.
#pragma once
#include <cuda_runtime.h>
#include <string>
#include <vector>

namespace e55 {
namespace tc {

struct HeadWeights {
  const float* policy_w; const float* policy_b;   // host, folded [cin][69], [69]
  const float* d_value_w; float value_b;          // device [filters]
  const float* d_fc1_k; const float* d_fc1_b; const float* d_fc2_k; float fc2_b;  // device
};

struct State { int blocks = 0, filters = 0, in_planes = 0, sms = 0; };

inline void init(State& s, int blocks, int filters, int in_planes, int sms) { s.blocks = blocks; s.filters = filters; s.in_planes = in_planes; s.sms = sms; }
inline bool supported(const State&) { return false; }
inline void free_weights(State&) {}
inline void free_buffers(State&) {}
inline cudaError_t alloc_buffers(State&, int) { return cudaSuccess; }
inline cudaError_t set_weights(State&, const std::vector<std::vector<float>>&, const std::vector<std::vector<float>>&, const HeadWeights&) { return cudaSuccess; }
inline cudaError_t forward(State&, const float*, int, float*, float*, cudaStream_t, int*, std::string* why) { *why = "tensor-core path not built"; return cudaSuccess; }
inline cudaError_t export_trunk(State&, int, float*, cudaStream_t, int*, std::string* why) { *why = "tensor-core path not built"; return cudaSuccess; }

}  // namespace tc
}  // namespace e55

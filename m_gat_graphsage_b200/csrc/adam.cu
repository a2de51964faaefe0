// Adam update of train.py / ablation/model1.py (torch.optim.Adam: L2 weight decay folded into the gradient, bias-corrected
// moments, no amsgrad) as ONE launch over every parameter tensor of the model.
//
// Why our own: PyTorch's fused implementation (multi_tensor_apply) hands each CTA a 64 Ki-element chunk; the model1 trunk
// has 1.6 M parameters = 26 CTAs on 148 SMs, 80 us per step for 45 MB of traffic (0.56 TB/s).  Here a chunk is 1024
// elements (one float4 per thread), chunks are dealt round-robin to 4 CTAs per SM: the same 45 MB at HBM / L2 speed.
// Arithmetic in fp32, in the order of torch/optim/adam.py:_single_tensor_adam.
#include "common.cuh"

#include <cmath>

namespace mgs {
namespace {

constexpr int kAdamTensors = 24;                             // tensors per launch (kernel parameter space)
constexpr int kAdamChunk = 1024;

struct AdamTensor {
  float* p;
  const float* g;
  float* m;
  float* v;
  int64_t n;
  int64_t chunk0;                                            // first chunk id of this tensor
};
struct AdamArgs {
  AdamTensor t[kAdamTensors];
  int count;
  int64_t chunks;
};

// beta1 here is 1 - beta1 and the second beta2 argument 1 - beta2, both rounded from the DOUBLE differences as PyTorch does
// (1.f - 0.999f is 1.3e-5 away from 0.001f)
__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, float lr_bc1, float omb1, float beta2,
                                         float omb2, float eps, float wd, float bc2_sqrt) {
  if (wd != 0.f) g = fmaf(p, wd, g);                          // grad = grad.add(param, alpha=weight_decay)
  m = m + (g - m) * omb1;                                    // exp_avg.lerp_(grad, 1 - beta1)
  v = v * beta2 + (g * g) * omb2;                            // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
  const float denom = sqrtf(v) / bc2_sqrt + eps;             // (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
  p = p - lr_bc1 * (m / denom);                              // param.addcdiv_(exp_avg, denom, value=-step_size)
}

__global__ void __launch_bounds__(256) adam_kernel(const AdamArgs a, float lr_bc1, float omb1, float beta2, float omb2,
                                                   float eps, float wd, float bc2_sqrt) {
  for (int64_t c = blockIdx.x; c < a.chunks; c += gridDim.x) {
    int ti = 0;
#pragma unroll 1
    while (ti + 1 < a.count && a.t[ti + 1].chunk0 <= c) ++ti;
    const AdamTensor& t = a.t[ti];
    const int64_t i0 = (c - t.chunk0) * kAdamChunk + threadIdx.x * 4;
    if (i0 >= t.n) continue;
    const bool vec = i0 + 4 <= t.n && (((uintptr_t)t.p | (uintptr_t)t.g | (uintptr_t)t.m | (uintptr_t)t.v) & 15u) == 0;
    if (vec) {
      float4 p = *reinterpret_cast<const float4*>(t.p + i0), g = *reinterpret_cast<const float4*>(t.g + i0);
      float4 m = *reinterpret_cast<const float4*>(t.m + i0), v = *reinterpret_cast<const float4*>(t.v + i0);
      adam_one(p.x, g.x, m.x, v.x, lr_bc1, omb1, beta2, omb2, eps, wd, bc2_sqrt);
      adam_one(p.y, g.y, m.y, v.y, lr_bc1, omb1, beta2, omb2, eps, wd, bc2_sqrt);
      adam_one(p.z, g.z, m.z, v.z, lr_bc1, omb1, beta2, omb2, eps, wd, bc2_sqrt);
      adam_one(p.w, g.w, m.w, v.w, lr_bc1, omb1, beta2, omb2, eps, wd, bc2_sqrt);
      *reinterpret_cast<float4*>(t.p + i0) = p;
      *reinterpret_cast<float4*>(t.m + i0) = m;
      *reinterpret_cast<float4*>(t.v + i0) = v;
    } else {
      for (int64_t i = i0; i < min(i0 + 4, t.n); ++i) {
        float p = t.p[i], m = t.m[i], v = t.v[i];
        adam_one(p, t.g[i], m, v, lr_bc1, omb1, beta2, omb2, eps, wd, bc2_sqrt);
        t.p[i] = p; t.m[i] = m; t.v[i] = v;
      }
    }
  }
}

}  // namespace
}  // namespace mgs

using namespace mgs;

extern "C" int mgs_adam_step(int32_t count, float* const* params, const float* const* grads, float* const* exp_avg,
                             float* const* exp_avg_sq, const int64_t* numel, double lr, double beta1, double beta2, double eps,
                             double weight_decay, int64_t step, mgs_stream_t stream_) {
  MGS_REQUIRE(count >= 0 && step >= 1, "mgs_adam_step: bad count / step");
  MGS_REQUIRE(count == 0 || (params && grads && exp_avg && exp_avg_sq && numel), "mgs_adam_step: null table");
  cudaStream_t stream = (cudaStream_t)stream_;
  const double bc1 = 1.0 - std::pow(beta1, (double)step);
  const double bc2 = 1.0 - std::pow(beta2, (double)step);
  const float lr_bc1 = (float)(lr / bc1);
  const float bc2_sqrt = (float)std::sqrt(bc2);
  for (int base = 0; base < count; base += kAdamTensors) {
    AdamArgs a = {};
    int64_t chunks = 0;
    for (int i = base; i < count && a.count < kAdamTensors; ++i) {
      if (numel[i] <= 0) continue;
      MGS_REQUIRE(params[i] && grads[i] && exp_avg[i] && exp_avg_sq[i], "mgs_adam_step: null tensor %d", i);
      AdamTensor& t = a.t[a.count++];
      t.p = params[i]; t.g = grads[i]; t.m = exp_avg[i]; t.v = exp_avg_sq[i];
      t.n = numel[i];
      t.chunk0 = chunks;
      chunks += (numel[i] + kAdamChunk - 1) / kAdamChunk;
    }
    if (a.count == 0) continue;
    a.chunks = chunks;
    const int grid = grid_for(chunks * 256, 256, 4);
    adam_kernel<<<grid, 256, 0, stream>>>(a, lr_bc1, (float)(1.0 - beta1), (float)beta2, (float)(1.0 - beta2), (float)eps,
                                          (float)weight_decay, bc2_sqrt);
    if (int rc = check_launch("adam_kernel")) return rc;
  }
  return MGS_OK;
}

// Device image preprocessing: the image processor's resize + rescale + normalise (reference
// detikzify/model/v1/processing_detikzify.py:242-251: PIL bicubic resize to S x S, x 1/255, (x - mean) / std, CHW) for
// candidate renders that arrive in bursts from parallel MCTS rollouts. The resize reproduces Pillow's 8-bit resampler
// bit for bit (Pillow src/libImaging/Resample.c, third party: separable convolution, horizontal pass then vertical
// pass, 22-bit fixed-point coefficients, rounding to uint8 after each pass); the coefficient tables depend only on
// (input size, output size) and are computed on the host by the Python shim (model/processing.py::pil_resample_coeffs).
// HBM-bound byte work: one thread per output pixel (3 channels), coalesced uint8 rows.
#include "common.cuh"
#include "launch.h"

namespace dtk {
namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;

DTK_DEV uint8_t clip8(int v) {
  v >>= PRECISION_BITS;
  return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// in [H, Win, 3] -> out [H, Wout, 3]; bounds[x] = {xmin, xcount}, coef [Wout][ksize]
__global__ void resample_h_kernel(const uint8_t* __restrict__ in, int H, int Win, int Wout, const int* __restrict__ bounds,
                                  const int* __restrict__ coef, int ksize, uint8_t* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= Wout) return;
  const int xmin = bounds[2 * x], n = bounds[2 * x + 1];
  const int* k = coef + (int64_t)x * ksize;
  const uint8_t* row = in + ((int64_t)y * Win + xmin) * 3;
  int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
  for (int i = 0; i < n; ++i) {
    const int kk = k[i];
    s0 += row[3 * i] * kk; s1 += row[3 * i + 1] * kk; s2 += row[3 * i + 2] * kk;
  }
  uint8_t* o = out + ((int64_t)y * Wout + x) * 3;
  o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
}

// in [Hin, W, 3] (uint8) -> out fp32 [3, Hout, W] normalised: (v / 255 - mean) / std
__global__ void resample_v_norm_kernel(const uint8_t* __restrict__ in, int Hin, int W, int Hout, const int* __restrict__ bounds,
                                       const int* __restrict__ coef, int ksize, float rescale, float3 mean, float3 istd,
                                       float* __restrict__ out, uint8_t* __restrict__ out_u8) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W) return;
  const int ymin = bounds[2 * y], n = bounds[2 * y + 1];
  const int* k = coef + (int64_t)y * ksize;
  int s0 = 1 << (PRECISION_BITS - 1), s1 = s0, s2 = s0;
  for (int i = 0; i < n; ++i) {
    const uint8_t* px = in + ((int64_t)(ymin + i) * W + x) * 3;
    const int kk = k[i];
    s0 += px[0] * kk; s1 += px[1] * kk; s2 += px[2] * kk;
  }
  const uint8_t c0 = clip8(s0), c1 = clip8(s1), c2 = clip8(s2);
  if (out_u8) {
    uint8_t* o = out_u8 + ((int64_t)y * W + x) * 3;
    o[0] = c0; o[1] = c1; o[2] = c2;
  }
  const int64_t plane = (int64_t)Hout * W, at = (int64_t)y * W + x;
  out[at] = ((float)c0 * rescale - mean.x) * istd.x;
  out[plane + at] = ((float)c1 * rescale - mean.y) * istd.y;
  out[2 * plane + at] = ((float)c2 * rescale - mean.z) * istd.z;
}

}  // namespace

cudaError_t launch_image_preprocess(const uint8_t* rgb, int h, int w, int S, const int* bounds_h, const int* coef_h, int ksize_h,
                                    const int* bounds_v, const int* coef_v, int ksize_v, float rescale, const float* mean,
                                    const float* std, uint8_t* tmp, float* out, uint8_t* out_u8, cudaStream_t s, uint64_t* counter) {
  if (h <= 0 || w <= 0 || S <= 0) return cudaErrorInvalidValue;
  dim3 gh((S + 127) / 128, h), gv((S + 127) / 128, S);
  resample_h_kernel<<<gh, 128, 0, s>>>(rgb, h, w, S, bounds_h, coef_h, ksize_h, tmp);
  resample_v_norm_kernel<<<gv, 128, 0, s>>>(tmp, h, S, S, bounds_v, coef_v, ksize_v, rescale, make_float3(mean[0], mean[1], mean[2]),
                                            make_float3(1.f / std[0], 1.f / std[1], 1.f / std[2]), out, out_u8);
  if (counter) *counter += 2;
  return cudaGetLastError();
}

}  // namespace dtk

// Shared device helpers for the encode / overlap kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace oc {

constexpr unsigned FULL = 0xffffffffu;

// Per-device "already done" flags (kernel attributes such as the > 48 KB shared-memory opt-in are
// per device: a process that uses a second GPU must set them there too).
constexpr int kMaxDevices = 64;
struct PerDeviceFlag {
  bool done[kMaxDevices] = {};
  // returns the flag of the calling thread's current device
  bool& cur() {
    int dev = 0;
    cudaGetDevice(&dev);
    return done[(dev >= 0 && dev < kMaxDevices) ? dev : 0];
  }
};

template <typename T> struct Num;
template <> struct Num<float> {
  // rotate a pair only if |a.b| > tol * |a||b|  (relative, Hestenes/de Rijk)
  static __device__ __forceinline__ float tol() { return 1e-6f; }
  // vectors below floor * ||M||_F are numerically zero (noise of earlier rotations)
  static __device__ __forceinline__ float floor_rel() { return 1e-6f; }
  static __device__ __forceinline__ float rsqrt(float x) { return rsqrtf(x); }
  static __device__ __forceinline__ float sqrt(float x) { return sqrtf(x); }
  static __device__ __forceinline__ float abs(float x) { return fabsf(x); }
};
template <> struct Num<double> {
  static __device__ __forceinline__ double tol() { return 1e-14; }
  static __device__ __forceinline__ double floor_rel() { return 1e-14; }
  static __device__ __forceinline__ double rsqrt(double x) { return 1.0 / ::sqrt(x); }
  static __device__ __forceinline__ double sqrt(double x) { return ::sqrt(x); }
  static __device__ __forceinline__ double abs(double x) { return fabs(x); }
};

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

// Jacobi rotation for a pair with Gram entries aa, bb, ab:
//   a' = c a - s b,  b' = s a + c b  are orthogonal.  |theta| <= pi/4.
template <typename T>
__device__ __forceinline__ void jacobi_cs(T aa, T bb, T ab, T& c, T& s) {
  T zeta = (bb - aa) / (T(2) * ab);
  T az = Num<T>::abs(zeta);
  T t = T(1) / (az + Num<T>::sqrt(T(1) + zeta * zeta));
  t = zeta < T(0) ? -t : t;
  c = Num<T>::rsqrt(T(1) + t * t);
  s = c * t;
}

template <typename T>
__device__ __forceinline__ T load_pixel(const void* img, int in_dtype, int64_t i) {
  if (in_dtype == 0) return T(reinterpret_cast<const float*>(img)[i]);
  if (in_dtype == 1) return T(reinterpret_cast<const double*>(img)[i]);
  return T(reinterpret_cast<const uint8_t*>(img)[i]) / T(255);
}

}  // namespace oc

// RMSNorm / fused residual-add RMSNorm / LayerNorm / gated activations (sm_100a).
//
// Reference ops: csrc/libtorch_stable/layernorm_kernels.cu (rms_norm,
// fused_add_rms_norm), activation_kernels.cu (silu_and_mul, gelu_and_mul, ...).
// Our variants: one CTA per token, 16-byte vector access, the row stays in
// registers between the reduction and the scale pass (single HBM read), fp32 math.
#include "common.cuh"

namespace {

constexpr int MAXV = 4;  // 16B vectors per thread kept in registers

template <typename T, bool ADD, bool GEMMA>
__global__ void __launch_bounds__(1024) rms_norm_kernel(T* __restrict__ out, T* __restrict__ x,
                                                         T* __restrict__ residual, const T* __restrict__ w,
                                                         int H, float eps, int64_t x_stride) {
  constexpr int N = Vec16<T>::N;
  __shared__ float red[32];
  pdl_trigger();
  pdl_wait();
  const int64_t row = blockIdx.x;
  T* xr = x + row * x_stride;
  T* rr = ADD ? residual + row * (int64_t)H : nullptr;
  T* orow = out + row * (ADD ? x_stride : (int64_t)H);
  const int nvec = H / N;
  Vec16<T> v[MAXV];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = threadIdx.x + i * blockDim.x;
    if (vi < nvec) {
      v[i] = ld16(xr + vi * N);
      if (ADD) {
        Vec16<T> r = ld16(rr + vi * N);
#pragma unroll
        for (int j = 0; j < N; ++j) v[i].e[j] = from_f32<T>(to_f32(v[i].e[j]) + to_f32(r.e[j]));
        st16(rr + vi * N, v[i]);
      }
#pragma unroll
      for (int j = 0; j < N; ++j) { float f = to_f32(v[i].e[j]); ss += f * f; }
    }
  }
  ss = block_sum(ss, red);
  const float inv = rsqrtf(ss / (float)H + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = threadIdx.x + i * blockDim.x;
    if (vi < nvec) {
      Vec16<T> wv = ld16(w + vi * N), o;
#pragma unroll
      for (int j = 0; j < N; ++j) {
        float g = to_f32(wv.e[j]);
        if (GEMMA) g += 1.f;
        o.e[j] = from_f32<T>(to_f32(v[i].e[j]) * inv * g);
      }
      st16(orow + vi * N, o);
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(1024) layer_norm_kernel(T* __restrict__ out, const T* __restrict__ x,
                                                           const T* __restrict__ w, const T* __restrict__ b, int H,
                                                           float eps) {
  constexpr int N = Vec16<T>::N;
  __shared__ float red[32];
  const int64_t row = blockIdx.x;
  const T* xr = x + row * H;
  const int nvec = H / N;
  Vec16<T> v[MAXV];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = threadIdx.x + i * blockDim.x;
    if (vi < nvec) {
      v[i] = ld16(xr + vi * N);
#pragma unroll
      for (int j = 0; j < N; ++j) s += to_f32(v[i].e[j]);
    }
  }
  const float mean = block_sum(s, red) / (float)H;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = threadIdx.x + i * blockDim.x;
    if (vi < nvec) {
#pragma unroll
      for (int j = 0; j < N; ++j) { float d = to_f32(v[i].e[j]) - mean; ss += d * d; }
    }
  }
  const float inv = rsqrtf(block_sum(ss, red) / (float)H + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int vi = threadIdx.x + i * blockDim.x;
    if (vi < nvec) {
      Vec16<T> wv = ld16(w + vi * N), bv, o;
      if (b) bv = ld16(b + vi * N);
#pragma unroll
      for (int j = 0; j < N; ++j) {
        float y = (to_f32(v[i].e[j]) - mean) * inv * to_f32(wv.e[j]);
        if (b) y += to_f32(bv.e[j]);
        o.e[j] = from_f32<T>(y);
      }
      st16(out + row * H + vi * N, o);
    }
  }
}

enum : int { ACT_SILU = 0, ACT_GELU = 1, ACT_GELU_TANH = 2, ACT_RELU = 3, ACT_GELU_NEW = 4, ACT_GELU_FAST = 5,
             ACT_QUICK_GELU = 6, ACT_RELU2 = 7 };

template <int ACT> VB_DEVICE float act_fn(float x) {
  if (ACT == ACT_SILU) return x / (1.f + __expf(-x));
  if (ACT == ACT_GELU) return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
  if (ACT == ACT_GELU_TANH || ACT == ACT_GELU_NEW)
    return 0.5f * x * (1.f + tanhf(0.7978845608028654f * (x + 0.044715f * x * x * x)));
  if (ACT == ACT_GELU_FAST) return 0.5f * x * (1.f + tanhf(x * 0.7978845608f * (1.f + 0.044715f * x * x)));
  if (ACT == ACT_QUICK_GELU) return x / (1.f + __expf(-1.702f * x));
  if (ACT == ACT_RELU2) { float r = fmaxf(x, 0.f); return r * r; }
  return fmaxf(x, 0.f);
}

// out[t, :d] = act(x[t, :d]) * x[t, d:2d]
template <typename T, int ACT>
__global__ void act_and_mul_kernel(T* __restrict__ out, const T* __restrict__ x, int d, int64_t total_vec) {
  constexpr int N = Vec16<T>::N;
  pdl_trigger();
  pdl_wait();
  const int vec_per_row = d / N;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i / vec_per_row;
    const int c = (int)(i - t * vec_per_row) * N;
    Vec16<T> a = ld16(x + t * 2 * d + c), b = ld16(x + t * 2 * d + d + c), o;
#pragma unroll
    for (int j = 0; j < N; ++j) o.e[j] = from_f32<T>(act_fn<ACT>(to_f32(a.e[j])) * to_f32(b.e[j]));
    st16(out + t * d + c, o);
  }
}

// interleaved variant: x[t, 2i] = gate_i, x[t, 2i+1] = up_i (the weight layout of the fused gate_up GEMM epilogue)
template <typename T>
__global__ void silu_mul_interleaved_kernel(T* __restrict__ out, const T* __restrict__ x, int d, int64_t total_vec) {
  constexpr int N = Vec16<T>::N;
  pdl_trigger();
  pdl_wait();
  const int vec_per_row = d / N;                       // output vectors per row; each consumes two input vectors
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t t = i / vec_per_row;
    const int c = (int)(i - t * vec_per_row) * N;
    Vec16<T> a = ld16(x + t * 2 * d + 2 * c), b = ld16(x + t * 2 * d + 2 * c + N), o;
#pragma unroll
    for (int j = 0; j < N / 2; ++j) {
      o.e[j] = from_f32<T>(act_fn<ACT_SILU>(to_f32(a.e[2 * j])) * to_f32(a.e[2 * j + 1]));
      o.e[N / 2 + j] = from_f32<T>(act_fn<ACT_SILU>(to_f32(b.e[2 * j])) * to_f32(b.e[2 * j + 1]));
    }
    st16(out + t * d + c, o);
  }
}

template <typename T, int ACT>
__global__ void act_kernel(T* __restrict__ out, const T* __restrict__ x, int64_t total_vec) {
  constexpr int N = Vec16<T>::N;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total_vec; i += (int64_t)gridDim.x * blockDim.x) {
    Vec16<T> a = ld16(x + i * N), o;
#pragma unroll
    for (int j = 0; j < N; ++j) o.e[j] = from_f32<T>(act_fn<ACT>(to_f32(a.e[j])));
    st16(out + i * N, o);
  }
}

int norm_threads(int H, int vecN) {
  int nvec = H / vecN;
  int t = (nvec + MAXV - 1) / MAXV;
  t = ((t + 31) / 32) * 32;
  if (t < 32) t = 32;
  // prefer more threads (1 vector each) while the row fits
  int t1 = ((nvec + 31) / 32) * 32;
  if (t1 <= 1024) t = t1;
  return t > 1024 ? 1024 : t;
}

}  // namespace

// x: [T, H] (row stride x_stride elements), w: [H]. ADD: residual [T, H] contiguous updated in place, normed → x.
EXPORT int vb_rms_norm(void* out, void* x, void* residual, const void* w, int T, int H, int64_t x_stride, float eps,
                       int gemma, int dtype, cudaStream_t st) {
  if (T == 0) return 0;
  DISPATCH_DTYPE(dtype, TT, {
    constexpr int N = Vec16<TT>::N;
    if (H % N != 0 || H / N > 1024 * MAXV) return -1;
    const int thr = norm_threads(H, N);
    if (residual) {
      if (gemma) vb_launch(rms_norm_kernel<TT, true, true>, T, thr, 0, st, (TT*)x, (TT*)x, (TT*)residual, (const TT*)w, H, eps, x_stride);
      else vb_launch(rms_norm_kernel<TT, true, false>, T, thr, 0, st, (TT*)x, (TT*)x, (TT*)residual, (const TT*)w, H, eps, x_stride);
    } else {
      if (gemma) vb_launch(rms_norm_kernel<TT, false, true>, T, thr, 0, st, (TT*)out, (TT*)x, (TT*)nullptr, (const TT*)w, H, eps, x_stride);
      else vb_launch(rms_norm_kernel<TT, false, false>, T, thr, 0, st, (TT*)out, (TT*)x, (TT*)nullptr, (const TT*)w, H, eps, x_stride);
    }
  });
  return (int)cudaGetLastError();
}

EXPORT int vb_layer_norm(void* out, const void* x, const void* w, const void* b, int T, int H, float eps, int dtype,
                         cudaStream_t st) {
  if (T == 0) return 0;
  DISPATCH_DTYPE(dtype, TT, {
    constexpr int N = Vec16<TT>::N;
    if (H % N != 0 || H / N > 1024 * MAXV) return -1;
    layer_norm_kernel<TT><<<T, norm_threads(H, N), 0, st>>>((TT*)out, (const TT*)x, (const TT*)w, (const TT*)b, H, eps);
  });
  return (int)cudaGetLastError();
}

#define ACT_SWITCH(KERNEL, ...)                                                         \
  switch (act) {                                                                        \
    case ACT_SILU: KERNEL<TT, ACT_SILU> __VA_ARGS__; break;                             \
    case ACT_GELU: KERNEL<TT, ACT_GELU> __VA_ARGS__; break;                             \
    case ACT_GELU_TANH: KERNEL<TT, ACT_GELU_TANH> __VA_ARGS__; break;                   \
    case ACT_RELU: KERNEL<TT, ACT_RELU> __VA_ARGS__; break;                             \
    case ACT_GELU_NEW: KERNEL<TT, ACT_GELU_NEW> __VA_ARGS__; break;                     \
    case ACT_GELU_FAST: KERNEL<TT, ACT_GELU_FAST> __VA_ARGS__; break;                   \
    case ACT_QUICK_GELU: KERNEL<TT, ACT_QUICK_GELU> __VA_ARGS__; break;                 \
    default: KERNEL<TT, ACT_RELU2> __VA_ARGS__; break;                                  \
  }

EXPORT int vb_act_and_mul(void* out, const void* x, int64_t T, int d, int act, int dtype, cudaStream_t st) {
  if (T == 0) return 0;
  DISPATCH_DTYPE(dtype, TT, {
    constexpr int N = Vec16<TT>::N;
    if (d % N != 0) return -1;
    const int64_t total = T * (d / N);
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    switch (act) {
      case ACT_SILU: vb_launch(act_and_mul_kernel<TT, ACT_SILU>, blocks, 256, 0, st, (TT*)out, (const TT*)x, d, total); break;
      case ACT_GELU: vb_launch(act_and_mul_kernel<TT, ACT_GELU>, blocks, 256, 0, st, (TT*)out, (const TT*)x, d, total); break;
      default: vb_launch(act_and_mul_kernel<TT, ACT_GELU_TANH>, blocks, 256, 0, st, (TT*)out, (const TT*)x, d, total); break;
    }
  });
  return (int)cudaGetLastError();
}

EXPORT int vb_silu_mul_interleaved(void* out, const void* x, int64_t T, int d, int dtype, cudaStream_t st) {
  if (T == 0) return 0;
  if (dtype == DT_F32 || d % 8 != 0) return -1;
  const int64_t total = T * (d / 8);
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  DISPATCH_HALF(dtype, TT, {
    vb_launch(silu_mul_interleaved_kernel<TT>, blocks, 256, 0, st, (TT*)out, (const TT*)x, d, total);
  });
  return (int)cudaGetLastError();
}

EXPORT int vb_activation(void* out, const void* x, int64_t n, int act, int dtype, cudaStream_t st) {
  if (n == 0) return 0;
  DISPATCH_DTYPE(dtype, TT, {
    constexpr int N = Vec16<TT>::N;
    if (n % N != 0) return -1;
    const int64_t total = n / N;
    const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    ACT_SWITCH(act_kernel, <<<blocks, 256, 0, st>>>((TT*)out, (const TT*)x, total));
  });
  return (int)cudaGetLastError();
}

// microbench.cu — B200 pipe/atomic rates behind the kernel designs in DESIGN.md §4.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/microbench scripts/microbench.cu
// Prints one line per probe: warp-instructions (or lane-ops) per ns, chip-wide.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

constexpr int kIters = 2048;

__global__ void k_ffma(float* out, float a, float b) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = __fmaf_rn(x[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

__global__ void k_ffma2(float* out, float a, float b) {
  unsigned long long x[8];
  const float2 av = make_float2(a, a), bv = make_float2(b, b);
  const unsigned long long A = *reinterpret_cast<const unsigned long long*>(&av);
  const unsigned long long B = *reinterpret_cast<const unsigned long long*>(&bv);
#pragma unroll
  for (int i = 0; i < 8; ++i) { float2 v = make_float2(threadIdx.x + i, i); x[i] = *reinterpret_cast<unsigned long long*>(&v); }
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = ffma2(x[i], A, B);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) { float2 v = *reinterpret_cast<float2*>(&x[i]); s += v.x + v.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FFMA interleaved with integer ALU work (LOP3/IADD3): does the packed form free issue slots?
__global__ void k_mix(float* out, float a, float b, int packed) {
  unsigned long long x[4];
  float y[8];
  unsigned u[8];
  const float2 av = make_float2(a, a), bv = make_float2(b, b);
  const unsigned long long A = *reinterpret_cast<const unsigned long long*>(&av);
  const unsigned long long B = *reinterpret_cast<const unsigned long long*>(&bv);
#pragma unroll
  for (int i = 0; i < 8; ++i) { y[i] = threadIdx.x + i; u[i] = threadIdx.x * 7 + i; }
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 v = make_float2(threadIdx.x + i, i); x[i] = *reinterpret_cast<unsigned long long*>(&v); }
  if (packed) {
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
      for (int i = 0; i < 4; ++i) x[i] = ffma2(x[i], A, B);
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = (u[i] ^ (u[i] >> 3)) + 0x9e3779b9u;
    }
  } else {
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = __fmaf_rn(y[i], a, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) u[i] = (u[i] ^ (u[i] >> 3)) + 0x9e3779b9u;
    }
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += y[i] + (float)u[i];
#pragma unroll
  for (int i = 0; i < 4; ++i) { float2 v = *reinterpret_cast<float2*>(&x[i]); s += v.x + v.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_shfl(float* out) {
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] += __shfl_xor_sync(0xffffffffu, x[i], 1 + (it & 15));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// REDG patterns.  `faces` records of `stride` floats; every lane adds VEC floats per op.
// group = lanes of a warp that target the same face in one instruction (1 = all distinct).
template <int VEC>
__global__ void k_red(float* buf, int faces, int stride, int group, int ops, int active_lanes) {
  const int lane = threadIdx.x & 31;
  const unsigned gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (lane >= active_lanes) return;
  unsigned h = gw * 2654435761u;
  for (int i = 0; i < ops; ++i) {
    h = h * 1664525u + 1013904223u;
    // neighbouring warps / lanes hit neighbouring faces (screen locality)
    const unsigned f = ((h >> 8) + lane / group) % (unsigned)faces;
    float* p = buf + (size_t)f * stride + (VEC == 1 ? (i % 9) : 0);
    if (VEC == 1) atomicAdd(p, 1.0f);
    else if (VEC == 2) atomicAdd(reinterpret_cast<float2*>(p), make_float2(1.f, 2.f));
    else atomicAdd(reinterpret_cast<float4*>(p), make_float4(1.f, 2.f, 3.f, 4.f));
  }
}

template <typename F>
float time_ms(F launch, int reps = 5) {
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  launch();
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(a);
    launch();
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs\n", prop.name, sms);
  const int blocks = sms * 8, threads = 256;
  float* out;
  CK(cudaMalloc(&out, (size_t)blocks * threads * sizeof(float)));
  const double winstr = (double)blocks * threads / 32 * kIters * 8;
  {
    float ms = time_ms([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
    printf("FFMA   : %.1f warp-instr/ns chip (%.2f /clk/SM at 1.965 GHz)\n", winstr / ms / 1e6, winstr / ms / 1e6 / sms / 1.965);
    ms = time_ms([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
    printf("FFMA2  : %.1f warp-instr/ns chip (%.2f /clk/SM) = 2 FMAs each\n", winstr / ms / 1e6, winstr / ms / 1e6 / sms / 1.965);
    ms = time_ms([&] { k_shfl<<<blocks, threads>>>(out); });
    printf("SHFL+FADD: %.1f pairs/ns chip (%.2f /clk/SM)\n", winstr / ms / 1e6, winstr / ms / 1e6 / sms / 1.965);
    float m0 = time_ms([&] { k_mix<<<blocks, threads>>>(out, 1.0001f, 0.5f, 0); });
    float m1 = time_ms([&] { k_mix<<<blocks, threads>>>(out, 1.0001f, 0.5f, 1); });
    printf("mix 8 FMA + 16 ALU per iter: scalar %.3f ms, packed (4 FFMA2) %.3f ms\n", m0, m1);
  }
  // atomics
  const int faces = 655360;
  for (int stride : {16}) {
    float* buf;
    CK(cudaMalloc(&buf, (size_t)faces * stride * sizeof(float)));
    CK(cudaMemset(buf, 0, (size_t)faces * stride * sizeof(float)));
    const int ab = sms * 16, ops = 256;
    for (int group : {1, 4, 32}) {
      for (int lanes : {32, 8}) {
        const double laneops = (double)ab * threads / 32 * lanes * ops;
        float ms1 = time_ms([&] { k_red<1><<<ab, threads>>>(buf, faces, stride, group, ops, lanes); });
        float ms2 = time_ms([&] { k_red<2><<<ab, threads>>>(buf, faces, stride, group, ops, lanes); });
        float ms4 = time_ms([&] { k_red<4><<<ab, threads>>>(buf, faces, stride, group, ops, lanes); });
        printf("REDG stride %d group %2d lanes %2d: scalar %.1f, v2 %.1f, v4 %.1f lane-ops/ns chip\n", stride, group,
               lanes, laneops / ms1 / 1e6, laneops / ms2 / 1e6, laneops / ms4 / 1e6);
      }
    }
    cudaFree(buf);
  }
  return 0;
}

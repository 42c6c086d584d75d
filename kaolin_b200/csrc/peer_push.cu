// peer_push.cu — the exchange step of the view-sharded DIB-R path as a
// store-through-NVLink kernel: every rank PUSHES its per-view gradient shard into the gathered
// buffer of every peer (pointers into the peers' memory mapped in this process: CUDA symmetric /
// IPC memory, handed in by the host side, kaolin_b200/multi_gpu.py:PeerGradAllGather).
//
// An all-gather by stores needs no rendezvous inside the data path (a remote store is posted, a
// remote load would stall on the NVLink round trip) and no staging: one 16-byte load of the local
// shard (L2-resident: the backward just produced it) feeds one 16-byte store per destination.
// The NVSwitch gives every GPU its full 900 GB/s egress to any mix of peers, so the grid only
// has to keep enough 16-byte stores in flight; it is deliberately SMALL (a fraction of the SMs)
// because the kernel runs on a side stream underneath the soft-mask branch of the backward.
// Completion = kernel end (stores are performed at the destination before the stream moves on);
// the host side follows it with a cross-rank barrier before anyone reads a gathered buffer.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dibr_b200.h"

namespace {

constexpr int kMaxPeers = 16;
constexpr int kPushThreads = 512;
constexpr int kPushUnroll = 4;

struct PeerDst {
  int4* p[kMaxPeers];
};

__global__ void __launch_bounds__(kPushThreads) peer_push_kernel(const int4* __restrict__ src, size_t n16, PeerDst dst,
                                                                  int n_dst) {
  const size_t stride = (size_t)gridDim.x * kPushThreads;
  size_t i = (size_t)blockIdx.x * kPushThreads + threadIdx.x;
  // kPushUnroll independent 16-byte loads in flight per thread, then n_dst stores each
  for (; i + (kPushUnroll - 1) * stride < n16; i += kPushUnroll * stride) {
    int4 v[kPushUnroll];
#pragma unroll
    for (int u = 0; u < kPushUnroll; ++u) v[u] = __ldg(src + i + u * stride);
#pragma unroll
    for (int d = 0; d < kMaxPeers; ++d) {      // static indices: the pointers stay in the constant bank
      if (d < n_dst) {
#pragma unroll
        for (int u = 0; u < kPushUnroll; ++u) dst.p[d][i + u * stride] = v[u];
      }
    }
  }
  for (; i < n16; i += stride) {
    const int4 v = __ldg(src + i);
#pragma unroll
    for (int d = 0; d < kMaxPeers; ++d)
      if (d < n_dst) dst.p[d][i] = v;
  }
}

// Multicast variant: ONE multimem.st per 16 bytes to a multicast address - the NVSwitch replicates the
// store into the same offset of every GPU bound to the multicast object (this GPU included), so the
// egress of a rank is its shard once instead of once per peer (NVLink SHARP / "NVLS").
__device__ __forceinline__ void multimem_st_v4(void* mc, const int4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

__global__ void __launch_bounds__(kPushThreads) peer_push_multicast_kernel(const int4* __restrict__ src, size_t n16,
                                                                            int4* mc) {
  const size_t stride = (size_t)gridDim.x * kPushThreads;
  size_t i = (size_t)blockIdx.x * kPushThreads + threadIdx.x;
  for (; i + (kPushUnroll - 1) * stride < n16; i += kPushUnroll * stride) {
    int4 v[kPushUnroll];
#pragma unroll
    for (int u = 0; u < kPushUnroll; ++u) v[u] = __ldg(src + i + u * stride);
#pragma unroll
    for (int u = 0; u < kPushUnroll; ++u) multimem_st_v4(mc + i + u * stride, v[u]);
  }
  for (; i < n16; i += stride) multimem_st_v4(mc + i, __ldg(src + i));
}

}  // namespace

extern "C" int dibr_b200_peer_push_multicast(const void* src, size_t bytes, void* multicast_dst, size_t dst_offset_bytes,
                                             int ctas, dibr_b200_stream_t stream) {
  if (bytes && (!src || !multicast_dst)) return DIBR_B200_EINVAL;
  if ((bytes & 15) || (dst_offset_bytes & 15) || ((uintptr_t)src & 15) || ((uintptr_t)multicast_dst & 15)) return DIBR_B200_EINVAL;
  if (bytes == 0) return 0;
  const size_t n16 = bytes / 16;
  size_t want = (n16 + kPushThreads - 1) / kPushThreads;
  if (ctas <= 0) ctas = 32;
  if (want > (size_t)ctas) want = (size_t)ctas;
  peer_push_multicast_kernel<<<(unsigned)want, kPushThreads, 0, (cudaStream_t)stream>>>(
      static_cast<const int4*>(src), n16, reinterpret_cast<int4*>(static_cast<char*>(multicast_dst) + dst_offset_bytes));
  return (int)cudaGetLastError();
}

extern "C" int dibr_b200_peer_push(const void* src, size_t bytes, void* const* dst, int n_dst, size_t dst_offset_bytes,
                                   int ctas, dibr_b200_stream_t stream) {
  if (n_dst < 0 || n_dst > kMaxPeers || (bytes && (!src || (n_dst && !dst)))) return DIBR_B200_EINVAL;
  if ((bytes & 15) || (dst_offset_bytes & 15) || ((uintptr_t)src & 15)) return DIBR_B200_EINVAL;
  if (bytes == 0 || n_dst == 0) return 0;
  PeerDst d;
  for (int i = 0; i < n_dst; ++i) {
    if (!dst[i] || ((uintptr_t)dst[i] & 15)) return DIBR_B200_EINVAL;
    d.p[i] = reinterpret_cast<int4*>(static_cast<char*>(dst[i]) + dst_offset_bytes);
  }
  const size_t n16 = bytes / 16;
  size_t want = (n16 + kPushThreads - 1) / kPushThreads;
  if (ctas <= 0) ctas = 32;
  if (want > (size_t)ctas) want = (size_t)ctas;
  peer_push_kernel<<<(unsigned)want, kPushThreads, 0, (cudaStream_t)stream>>>(static_cast<const int4*>(src), n16, d, n_dst);
  return (int)cudaGetLastError();
}

// SyncBatchNorm exchange fused with its merge, over NVLink / NVSwitch PEER MEMORY (reference
// models/model_ctc.py:70-75 convert_sync_batchnorm: every synchronised BatchNorm layer exchanges [2C + 1] floats forward and [2C]
// floats backward -- 32 latency-bound collectives per training step of CTCSmall).
//
// One kernel per exchange instead of {pack, NCCL all_gather / all_reduce, unpack, merge}: every rank owns a MAILBOX in symmetric
// memory (allocated and peer-mapped by torch.distributed._symmetric_memory; this kernel only sees the mapped base pointers).  A rank
//   1. stores its payload into slot [parity][rank] of EVERY rank's mailbox (plain stores to peer addresses: they travel over NVLink),
//   2. fences (system scope) and raises flag [parity][rank] in every mailbox to the tag of this exchange (st.release.sys),
//   3. waits until all flags of its OWN mailbox carry the tag (ld.acquire.sys; the flags are written remotely, read locally),
//   4. merges the W payloads of its own mailbox in RANK ORDER (so every rank computes bit-identical results):
//        mode 0  element-wise sum                         (BatchNorm backward: sum dz, sum dz * xhat)
//        mode 1  Chan merge of (mean, M2) with counts     (BatchNorm forward statistics), total count to out_count
// The exchange epoch lives in the mailbox (device memory), so a captured CUDA graph replays correctly.  Two parities suffice:
// every exchange is a full barrier, so when a rank has finished exchange k every rank has finished reading exchange k-1, whose
// slots exchange k+1 reuses.  A rank that waits longer than ~4 s for a peer sets the error word and carries on (no GPU hang).
#include "ec_common.cuh"

namespace ec {

constexpr int kP2PMaxWorld = 16;
constexpr int kP2PMaxFloats = 4096;          // payload floats per rank and exchange (2C + 1 <= 4096)

struct P2PDev {
  float* buf[kP2PMaxWorld];                  // peer-mapped mailbox base of every rank, in this process's address space
  int rank, world;
};

__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
  unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}

// mailbox layout in floats: data [2][world][kP2PMaxFloats] | flags u32 [2][kP2PMaxWorld] | epoch u32 | error u32
__host__ __device__ inline size_t p2p_data_floats(int world) { return static_cast<size_t>(2) * world * kP2PMaxFloats; }

__global__ void __launch_bounds__(256) p2p_exchange_kernel(const P2PDev p, float* __restrict__ data, int n, int mode, float count,
                                                           float* __restrict__ out_count) {
  __shared__ unsigned s_epoch;
  const int tid = threadIdx.x, W = p.world, me = p.rank;
  const size_t dfl = p2p_data_floats(W);
  unsigned* ctl = reinterpret_cast<unsigned*>(p.buf[me] + dfl);          // own flags | epoch | error
  if (tid == 0) s_epoch = ctl[2 * kP2PMaxWorld];
  __syncthreads();
  const unsigned e = s_epoch, par = e & 1u, tag = e + 1u;
  // 1. payload -> slot [par][me] of every mailbox (own mailbox included)
  for (int r = 0; r < W; ++r) {
    float* dst = p.buf[r] + (static_cast<size_t>(par) * W + me) * kP2PMaxFloats;
    for (int i = tid; i < n; i += blockDim.x) dst[i] = data[i];
    if (mode == 1 && tid == 0) dst[n] = count;
  }
  __threadfence_system();
  __syncthreads();
  // 2. raise my flag everywhere;  3. wait for everybody's flag in my own mailbox
  if (tid < W) {
    unsigned* remote = reinterpret_cast<unsigned*>(p.buf[tid] + dfl) + par * kP2PMaxWorld + me;
    st_release_sys(remote, tag);
    const unsigned* mine = ctl + par * kP2PMaxWorld + tid;
    const long long t0 = clock64();
    while (ld_acquire_sys(mine) != tag) {
      if (clock64() - t0 > 8000000000LL) { ctl[2 * kP2PMaxWorld + 1] = 1u; break; }     // ~4 s at 2 GHz: report, do not hang
      __nanosleep(64);
    }
  }
  __syncthreads();
  // 4. merge in rank order
  const float* box = p.buf[me] + static_cast<size_t>(par) * W * kP2PMaxFloats;
  if (mode == 0) {
    for (int i = tid; i < n; i += blockDim.x) {
      float s = 0.f;
      for (int r = 0; r < W; ++r) s += box[static_cast<size_t>(r) * kP2PMaxFloats + i];
      data[i] = s;
    }
  } else {
    const int C = n / 2;
    for (int c = tid; c < C; c += blockDim.x) {
      double nn = 0.0, mean = 0.0, m2 = 0.0;
      for (int r = 0; r < W; ++r) {
        const float* b = box + static_cast<size_t>(r) * kP2PMaxFloats;
        const double nr = b[n];
        if (nr <= 0.0) continue;
        const double tot = nn + nr, delta = static_cast<double>(b[c]) - mean;
        m2 += static_cast<double>(b[C + c]) + delta * delta * nn * nr / tot;
        mean += delta * nr / tot;
        nn = tot;
      }
      data[c] = static_cast<float>(mean);
      data[C + c] = static_cast<float>(m2);
    }
    if (tid == 0 && out_count != nullptr) {
      float tot = 0.f;
      for (int r = 0; r < W; ++r) tot += box[static_cast<size_t>(r) * kP2PMaxFloats + n];
      out_count[0] = tot;
    }
  }
  __syncthreads();
  if (tid == 0) ctl[2 * kP2PMaxWorld] = tag;
}

}  // namespace ec

using namespace ec;
extern "C" {
size_t ec_p2p_mailbox_bytes(int world) { return (p2p_data_floats(world) + 2 * kP2PMaxWorld + 2) * sizeof(float); }
int ec_p2p_max_payload_floats() { return kP2PMaxFloats - 1; }
int ec_p2p_bn_exchange(const unsigned long long* peer_ptrs, int rank, int world, float* data, int n, int mode, float count,
                       float* out_count, void* stream) {
  EC_REQUIRE(peer_ptrs && data && world >= 1 && world <= kP2PMaxWorld && rank >= 0 && rank < world, "bad peer table");
  EC_REQUIRE(n > 0 && n < kP2PMaxFloats && (mode == 0 || (mode == 1 && n % 2 == 0)), "bad payload");
  P2PDev p{};
  for (int r = 0; r < world; ++r) p.buf[r] = reinterpret_cast<float*>(static_cast<uintptr_t>(peer_ptrs[r]));
  p.rank = rank; p.world = world;
  p2p_exchange_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(p, data, n, mode, count, out_count);
  EC_CUDA(cudaGetLastError());
  return EC_OK;
}
/* error word of the mailbox (1 = a peer did not arrive within the time-out); synchronises the stream */
int ec_p2p_error(const unsigned long long* peer_ptrs, int rank, int world, int* out) {
  EC_REQUIRE(peer_ptrs && out, "null argument");
  const unsigned* ctl = reinterpret_cast<const unsigned*>(reinterpret_cast<const float*>(static_cast<uintptr_t>(peer_ptrs[rank])) + p2p_data_floats(world));
  unsigned v = 0;
  EC_CUDA(cudaMemcpy(&v, ctl + 2 * kP2PMaxWorld + 1, sizeof(unsigned), cudaMemcpyDeviceToHost));
  *out = static_cast<int>(v);
  return EC_OK;
}
}

// Shared pieces of the tcgen05 GEMM family (gemm_tc.cu, ffn_fused.cu): epilogue slab layouts matching the TMA swizzle modes,
// bulk-tensor store helpers, chunk-wise LayerNorm statistics and the host-side tensor-map builders.
#pragma once
#include "ec_common.cuh"
#include <mutex>
#include <type_traits>

namespace ec {

constexpr int kBlockM = 128;
constexpr int kATileBytes = kBlockM * 128;
constexpr int kSlabBytes = 4096;             // 32 rows x 128 B

// Mean / centred sum of squares of the first nc (<= 32) values of t; the caller guarantees t[j] == 0 for j >= nc, so the sums
// run over all 32 registers without per-element predicates and the zero tail is taken out in closed form
// (sum over the tail of (0 - mean)^2 = (32 - nc) mean^2).  Four independent accumulation chains.
__device__ __forceinline__ void chunk_stats(const float (&t)[32], int nc, float& cm, float& cq) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
  for (int j = 0; j < 32; j += 4) { s0 += t[j]; s1 += t[j + 1]; s2 += t[j + 2]; s3 += t[j + 3]; }
  cm = __fdividef((s0 + s1) + (s2 + s3), static_cast<float>(nc));
  float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f;
#pragma unroll
  for (int j = 0; j < 32; j += 4) {
    const float d0 = t[j] - cm, d1 = t[j + 1] - cm, d2 = t[j + 2] - cm, d3 = t[j + 3] - cm;
    q0 = fmaf(d0, d0, q0); q1 = fmaf(d1, d1, q1); q2 = fmaf(d2, d2, q2); q3 = fmaf(d3, d3, q3);
  }
  cq = fmaf(-static_cast<float>(32 - nc) * cm, cm, (q0 + q1) + (q2 + q3));
}
// Chan et al. merge of a chunk (nc values, mean cm, centred squares cq) into running (cnt, mean, m2).
__device__ __forceinline__ void stats_merge(float& cnt, float& mean, float& m2, float nc, float cm, float cq) {
  const float tot = cnt + nc, w = __fdividef(nc, tot), dlt = cm - mean;
  mean = fmaf(dlt, w, mean);
  m2 += fmaf(dlt * dlt, cnt * w, cq);
  cnt = tot;
}
// wait until at most `pending` (1, 3 or 7) of this thread's bulk-store groups have not finished reading shared memory
__device__ __forceinline__ void bulk_wait_read(int pending) {
  if (pending >= 7) asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory");
  else if (pending >= 3) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
  else asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}

// ---- TMA store / bulk-group helpers -------------------------------------------------------------------------------
__device__ __forceinline__ void tma_store_2d(const void* tmap, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// 32 x 32 slab layouts written by thread = row (lane), matching the TMA swizzle modes:
//   fp32: 128-byte rows, 16-byte chunk j4 in [0,8) XOR (row & 7)          (CU_TENSOR_MAP_SWIZZLE_128B)
//   bf16:  64-byte rows, 16-byte chunk j8 in [0,4) XOR ((row >> 1) & 3)   (CU_TENSOR_MAP_SWIZZLE_64B)
__device__ __forceinline__ uint32_t slab_f32_off(int row, int j4) { return row * 128 + ((j4 ^ (row & 7)) << 4); }
__device__ __forceinline__ uint32_t slab_b16_off(int row, int j8) { return row * 64 + ((j8 ^ ((row >> 1) & 3)) << 4); }

__device__ __forceinline__ void slab_store_f32(uint8_t* slab, int row, const float (&t)[32]) {
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4)
    *reinterpret_cast<float4*>(slab + slab_f32_off(row, j4)) = make_float4(t[4 * j4], t[4 * j4 + 1], t[4 * j4 + 2], t[4 * j4 + 3]);
}
__device__ __forceinline__ void slab_load_f32(const uint8_t* slab, int row, float (&t)[32]) {
#pragma unroll
  for (int j4 = 0; j4 < 8; ++j4) {
    const float4 x = *reinterpret_cast<const float4*>(slab + slab_f32_off(row, j4));
    t[4 * j4] = x.x; t[4 * j4 + 1] = x.y; t[4 * j4 + 2] = x.z; t[4 * j4 + 3] = x.w;
  }
}
template <typename T>
__device__ __forceinline__ void slab_store_act(uint8_t* slab, int row, const float (&t)[32]) {
  if constexpr (IsSplit<T>::value) {
    float r[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = __uint_as_float(split_pack(t[j]));
    slab_store_f32(slab, row, r);
  } else if constexpr (sizeof(T) == 4) {
    float r[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = round_tf32(t[j]);
    slab_store_f32(slab, row, r);
  } else if constexpr (sizeof(T) == 2 && !std::is_same<T, __nv_bfloat16>::value) {     // fp16
#pragma unroll
    for (int j8 = 0; j8 < 4; ++j8) {
      uint4 pk;
      __half2 h0 = __floats2half2_rn(t[8 * j8], t[8 * j8 + 1]), h1 = __floats2half2_rn(t[8 * j8 + 2], t[8 * j8 + 3]);
      __half2 h2 = __floats2half2_rn(t[8 * j8 + 4], t[8 * j8 + 5]), h3 = __floats2half2_rn(t[8 * j8 + 6], t[8 * j8 + 7]);
      pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
      pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
      *reinterpret_cast<uint4*>(slab + slab_b16_off(row, j8)) = pk;
    }
  } else {
#pragma unroll
    for (int j8 = 0; j8 < 4; ++j8) {
      uint4 pk;
      __nv_bfloat162 h0 = __floats2bfloat162_rn(t[8 * j8], t[8 * j8 + 1]), h1 = __floats2bfloat162_rn(t[8 * j8 + 2], t[8 * j8 + 3]);
      __nv_bfloat162 h2 = __floats2bfloat162_rn(t[8 * j8 + 4], t[8 * j8 + 5]), h3 = __floats2bfloat162_rn(t[8 * j8 + 6], t[8 * j8 + 7]);
      pk.x = *reinterpret_cast<uint32_t*>(&h0); pk.y = *reinterpret_cast<uint32_t*>(&h1);
      pk.z = *reinterpret_cast<uint32_t*>(&h2); pk.w = *reinterpret_cast<uint32_t*>(&h3);
      *reinterpret_cast<uint4*>(slab + slab_b16_off(row, j8)) = pk;
    }
  }
}

// fp32 values of a 32 x 32 activation-type slab loaded by TMA (packed / TF32 words: SW128 fp32 layout; bf16: SW64 layout)
template <typename T>
__device__ __forceinline__ void slab_load_act(const uint8_t* slab, int row, float (&t)[32]) {
  if constexpr (sizeof(T) == 4) {
    slab_load_f32(slab, row, t);
    if constexpr (IsSplit<T>::value) {
#pragma unroll
      for (int j = 0; j < 32; ++j) t[j] = split_unpack(__float_as_uint(t[j]));
    }
  } else {
#pragma unroll
    for (int j8 = 0; j8 < 4; ++j8) {
      const uint4 pk = *reinterpret_cast<const uint4*>(slab + slab_b16_off(row, j8));
      const uint32_t w[4] = {pk.x, pk.y, pk.z, pk.w};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        t[8 * j8 + 2 * k] = __uint_as_float(w[k] << 16);
        t[8 * j8 + 2 * k + 1] = __uint_as_float(w[k] & 0xffff0000u);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// host side: tensor maps
// ------------------------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2D row-major tensor [rows, cols] of `esize`-byte elements -> tensor map with a (box_cols x box_rows) box.
inline int make_map(CUtensorMap* map, bool is_f32, const void* ptr, int rows, int cols, int ld, int box_cols, int box_rows,
                    CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = get_encode_fn();
  EC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled is not available from the CUDA driver");
  const int esize = is_f32 ? 4 : 2;
  const size_t pitch = static_cast<size_t>(ld) * esize;
  EC_REQUIRE(pitch % 16 == 0, "tensor row pitch must be a multiple of 16 bytes for TMA (got " + std::to_string(pitch) + ")");
  EC_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "tensor base must be 16-byte aligned for TMA");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(pitch)};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, is_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims,
                   strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  EC_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with code " + std::to_string(static_cast<int>(r)));
  return EC_OK;
}
// K-major GEMM operand [rows, K]: (128-byte x box_rows) box, 128B swizzle.
inline int make_operand_map(CUtensorMap* map, int precision, const void* ptr, int rows, int K, int box_rows) {
  const bool f32 = precision != EC_PREC_BF16;      // split mode: 4-byte packed pairs move like fp32 words
  return make_map(map, f32, ptr, rows, K, K, f32 ? 32 : 64, box_rows, CU_TENSOR_MAP_SWIZZLE_128B);
}
// 32 x 32 epilogue slab of an [M, cols] output / residual tensor.
inline int make_slab_map(CUtensorMap* map, bool is_f32, const void* ptr, int rows, int cols, int ld) {
  return make_map(map, is_f32, ptr, rows, cols, ld, 32, 32, is_f32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
}

}  // namespace ec

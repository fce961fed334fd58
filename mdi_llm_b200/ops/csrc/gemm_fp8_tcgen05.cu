// Block-scaled fp8 (e4m3) GEMM on the 5th-generation tensor cores: tcgen05.mma kind::f8f6f4, sm_100a.
//
//   C[M, N] = sum_kb  a_s[m, kb] * w_s[n, kb] * ( A8[m, kb*128 : +128] . W8[n, kb*128 : +128] )   (+bias, +residual)
//
// The serving format of BASELINE config #5 ("fp8 block-scaled"): weights are e4m3 with one fp32 scale per 128
// consecutive K elements of every output row (what the decode kernels stream), activations are quantised the
// same way per token.  Scales change every 128 K elements, so the tensor core cannot accumulate the whole K
// range in one go: every K block (= one 128-byte shared-memory slab = 4 MMAs of K = 32) is accumulated into a
// FRESH partial accumulator in TMEM, and "promotion" warps fold it into fp32 register accumulators with the
// block's scales while the tensor core already works on the next blocks in the other TMEM buffers (4 x 128 columns):
//
//   warp 0      TMA producer: A8 / W8 tiles (128 rows x 128 B, SWIZZLE_128B) -> 5-stage shared-memory ring
//   warp 1      MMA issuer: per K block 4 x tcgen05.mma.kind::f8f6f4 (M=128, N=128, K=32) into one of 4 partial buffers,
//               tcgen05.commit -> "stage free" and "partial full"
//   warp 2      TMEM allocator
//   warps 4-11  promotion + epilogue: tcgen05.ld (warpgroup 0: columns 0-63, warpgroup 1: columns 64-127),
//               acc += partial * (a_s[row] * w_s[col]); after the last K block: bias / residual / activation,
//               bf16 stores (optionally straight into the next stage's memory: the fused prefill hop)
//
// Gated MLPs (SwiGLU / GeGLU) in one pass: the B tile is 64 rows of fc_1 and the SAME 64 rows of fc_2, so the
// output tile holds g in columns 0-63 and u in columns 64-127; warpgroup 1 hands u over through shared memory and
// warpgroup 0 writes act(g) * u.
#include <cuda_fp8.h>

#include "tcgen05.cuh"

namespace mdi {

constexpr int F8_BM = 128, F8_BN = 128, F8_BK = 128;  // BK in elements == bytes
constexpr int F8_STAGES = 5;
constexpr int F8_THREADS = 384;
constexpr int F8_TILE_BYTES = F8_BM * F8_BK;  // 16 KB (A) and 16 KB (B) per stage
// Block scales travel with the tiles: per K block 128 weight-column scales + 128 activation-row scales (2 x 512 B bulk
// copies onto the stage's "full" barrier) into a small ring of their own.  A scale slot outlives its tile stage — the
// promotion warps still read it after the MMAs have released the stage — so the ring is 3 slots deeper: stage reuse
// means MMA(it - STAGES) retired, hence promotion(it - STAGES - PBUF) has started, hence promotion(it - STAGES - PBUF - 1) is done.
constexpr int F8_PBUF = 4;  // partial accumulators in TMEM (4 x 128 columns = all 512): depth of the MMA <-> promotion pipeline
constexpr int F8_SCALE_SLOTS = F8_STAGES + F8_PBUF + 1;

// packed fp32 pair arithmetic (sm_100: FFMA2 / FMUL2 — two lanes per issue slot)
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ unsigned long long mul2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
// explicit shared-space loads: the pointers below derive from the re-aligned dynamic-smem base, which the compiler
// only knows as a generic address (it emitted 64-bit generic LD.E for them)
__device__ __forceinline__ void lds_v2x2(uint32_t saddr, unsigned long long& lo, unsigned long long& hi) {
  asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "r"(saddr));
}
__device__ __forceinline__ float lds_f32(uint32_t saddr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
}
__device__ __forceinline__ float lo32(unsigned long long v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi32(unsigned long long v) { return __uint_as_float((uint32_t)(v >> 32)); }

struct Fp8GemmParams {
  CUtensorMap tma_a;   // A8  [M, K] uint8, box 128 x 128
  CUtensorMap tma_b;   // W8  [N, K] uint8, box 128 x 128 (gated: box 64 x 128)
  CUtensorMap tma_b2;  // gated: second weight matrix, box 64 x 128
  const float* a_scale_t;   // [K/128, ld_as]  (transposed: a warp reads 32 consecutive rows' scales)
  const float* w_scale_t;   // [K/128, N]      (transposed copy of the decode kernels' [N, K/128])
  const float* w2_scale_t;  // gated
  long long ld_as;
  bf16* C;
  const bf16* bias;
  const bf16* bias2;
  const bf16* residual;
  int M, N, K;  // N = output columns (gated: the intermediate size)
  int act;      // gated: ACT_*_GATE; 0 = plain
  HopSignal signal;
  const int* ctx;
};

__device__ __forceinline__ void tcgen05_mma_f8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// cute::UMMA::InstrDescriptor for kind::f8f6f4: D = f32 (bits 4-5 = 1), A = B = E4M3 (format 0), both K-major
__device__ __forceinline__ constexpr uint32_t umma_idesc_e4m3(int m, int n) {
  return (1u << 4) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

template <bool GATED>
__global__ void __launch_bounds__(F8_THREADS, 1) gemm_fp8_blockscaled_kernel(const __grid_constant__ Fp8GemmParams p) {
  extern __shared__ __align__(1024) unsigned char f8_smem[];
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(f8_smem) + 1023) & ~uintptr_t(1023));
  unsigned char* smem_a = base;
  unsigned char* smem_b = base + F8_STAGES * F8_TILE_BYTES;
  float* sc_w = reinterpret_cast<float*>(smem_b + F8_STAGES * F8_TILE_BYTES);  // [F8_SCALE_SLOTS][128] weight-column scales
  float* sc_a = sc_w + F8_SCALE_SLOTS * 128;                                    // [F8_SCALE_SLOTS][128] activation-row scales
  float* xchg = sc_a + F8_SCALE_SLOTS * 128;  // gated: u hand-over [128 rows][64]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(xchg) + (GATED ? F8_BM * 64 * 4 : 0));
  uint64_t* empty_bar = full_bar + F8_STAGES;
  uint64_t* pfull_bar = empty_bar + F8_STAGES;   // [PBUF] partial accumulator complete
  uint64_t* pempty_bar = pfull_bar + F8_PBUF;    // [PBUF] partial accumulator drained (8 promotion warps)
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(pempty_bar + F8_PBUF);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_cols = GATED ? 64 : F8_BN;  // output columns per tile
  const int tiles_m = (p.M + F8_BM - 1) / F8_BM;
  const int tiles_n = (p.N + n_cols - 1) / n_cols;
  const int n_tiles = tiles_m * tiles_n;
  const int nkb = p.K / F8_BK;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tma_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tma_b) : "memory");
    if (GATED) asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tma_b2) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < F8_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < F8_PBUF; ++b) { mbar_init(&pfull_bar[b], 1); mbar_init(&pempty_bar[b], 8); }
    mbar_fence_init();
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(F8_PBUF * F8_BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = (tile % tiles_m) * F8_BM, n0 = (tile / tiles_m) * n_cols;
        for (int kb = 0; kb < nkb; ++kb, ++it) {
          const int s = it % F8_STAGES, sl = it % F8_SCALE_SLOTS;
          mbar_wait(&empty_bar[s], ((it / F8_STAGES) & 1) ^ 1);
          const int wn = min(n_cols, p.N - n0);  // valid output columns of this tile (N % 4 == 0: 16-byte copies)
          const uint32_t wbytes = (uint32_t)((GATED ? 2 : 1) * wn * 4);
          mbar_expect_tx(&full_bar[s], 2 * F8_TILE_BYTES + 512u + wbytes);
          bulk_g2s(sc_a + sl * 128, p.a_scale_t + (size_t)kb * p.ld_as + m0, 512u, &full_bar[s]);
          bulk_g2s(sc_w + sl * 128, p.w_scale_t + (size_t)kb * p.N + n0, (uint32_t)(wn * 4), &full_bar[s]);
          if (GATED) bulk_g2s(sc_w + sl * 128 + 64, p.w2_scale_t + (size_t)kb * p.N + n0, (uint32_t)(wn * 4), &full_bar[s]);
          tma_load_2d(smem_a + s * F8_TILE_BYTES, &p.tma_a, &full_bar[s], kb * F8_BK, m0);
          if (GATED) {  // rows 0-63 of the B tile: fc_1[n0 : n0+64], rows 64-127: fc_2[n0 : n0+64]
            tma_load_2d(smem_b + s * F8_TILE_BYTES, &p.tma_b, &full_bar[s], kb * F8_BK, n0);
            tma_load_2d(smem_b + s * F8_TILE_BYTES + F8_TILE_BYTES / 2, &p.tma_b2, &full_bar[s], kb * F8_BK, n0);
          } else {
            tma_load_2d(smem_b + s * F8_TILE_BYTES, &p.tma_b, &full_bar[s], kb * F8_BK, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one fresh partial accumulator per K block =====
    const uint32_t idesc = umma_idesc_e4m3(F8_BM, F8_BN);
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int s = it % F8_STAGES, pb = it % F8_PBUF;
        mbar_wait(&pempty_bar[pb], ((it / F8_PBUF) & 1) ^ 1);  // promotion warps are done with this buffer
        mbar_wait(&full_bar[s], (it / F8_STAGES) & 1);
        tcgen05_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(smem_a + s * F8_TILE_BYTES), b_addr = smem_u32(smem_b + s * F8_TILE_BYTES);
          const uint32_t tmem_acc = tmem_base + (uint32_t)(pb * F8_BN);
#pragma unroll
          for (int k = 0; k < F8_BK / 32; ++k)  // K = 32 elements = 32 bytes per MMA inside the 128-byte swizzle atom
            tcgen05_mma_f8(tmem_acc, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc, k ? 1u : 0u);
          tcgen05_commit(&empty_bar[s]);
          tcgen05_commit(&pfull_bar[pb]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===== promotion + epilogue =====
    const int pw = warp - 4, lg = pw & 3, half = pw >> 2;  // TMEM lanes [32*lg, +32), columns [64*half, +64)
    const bool vec_ok = (p.N % 8 == 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int m0 = (tile % tiles_m) * F8_BM, n0 = (tile / tiles_m) * n_cols;
      const int row = m0 + lg * 32 + lane;
      unsigned long long acc2[32];  // 64 fp32 accumulators as 32 packed pairs
#pragma unroll
      for (int j = 0; j < 32; ++j) acc2[j] = 0ull;
      for (int kb = 0; kb < nkb; ++kb, ++it) {
        const int pb = it % F8_PBUF, sl = it % F8_SCALE_SLOTS;
        mbar_wait(&pfull_bar[pb], (it / F8_PBUF) & 1);  // MMAs of this K block retired => its stage (tiles AND scales) had landed
        tcgen05_fence_after();
        const uint32_t taddr = tmem_base + (uint32_t)(pb * F8_BN + 64 * half) + ((uint32_t)(lg * 32) << 16);
        uint32_t t0[32], t1[32];
        tmem_ld_32x32(taddr, t0);
        tmem_ld_32x32(taddr + 32u, t1);
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&pempty_bar[pb]);  // the tensor core may overwrite this buffer
        const float a_s = lds_f32(smem_u32(sc_a + sl * 128 + lg * 32 + lane));
        const unsigned long long a2 = pack2(a_s, a_s);
        const uint32_t ws_addr = smem_u32(sc_w + sl * 128 + 64 * half);  // 64 column scales: broadcast 16-byte reads
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          unsigned long long w0a, w0b, w1a, w1b;
          lds_v2x2(ws_addr + 16 * j, w0a, w0b);
          lds_v2x2(ws_addr + 128 + 16 * j, w1a, w1b);
          acc2[2 * j] = fma2(pack2(__uint_as_float(t0[4 * j]), __uint_as_float(t0[4 * j + 1])), mul2(w0a, a2), acc2[2 * j]);
          acc2[2 * j + 1] = fma2(pack2(__uint_as_float(t0[4 * j + 2]), __uint_as_float(t0[4 * j + 3])), mul2(w0b, a2), acc2[2 * j + 1]);
          acc2[16 + 2 * j] = fma2(pack2(__uint_as_float(t1[4 * j]), __uint_as_float(t1[4 * j + 1])), mul2(w1a, a2), acc2[16 + 2 * j]);
          acc2[16 + 2 * j + 1] = fma2(pack2(__uint_as_float(t1[4 * j + 2]), __uint_as_float(t1[4 * j + 3])), mul2(w1b, a2), acc2[16 + 2 * j + 1]);
        }
      }
      float acc[64];
#pragma unroll
      for (int j = 0; j < 32; ++j) { acc[2 * j] = lo32(acc2[j]); acc[2 * j + 1] = hi32(acc2[j]); }
      // ---- epilogue ----
      if (GATED) {
        // warpgroup 1 holds u (fc_2), warpgroup 0 holds g (fc_1) for the same 64 output columns
        float* xrow = xchg + (size_t)(lg * 32 + lane) * 64;
        if (half == 1) {
#pragma unroll
          for (int j = 0; j < 64; ++j) {
            float u = acc[j];
            if (p.bias2 && n0 + j < p.N) u += __bfloat162float(p.bias2[n0 + j]);
            xrow[j] = round_bf16(u);
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 promotion warps only
        if (half == 0 && row < p.M) {
          bf16* crow = p.C + (size_t)row * p.N + n0;
#pragma unroll
          for (int v = 0; v < 8; ++v) {
            float f[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              float g = acc[v * 8 + j];
              if (p.bias && n0 + v * 8 + j < p.N) g += __bfloat162float(p.bias[n0 + v * 8 + j]);
              f[j] = round_bf16(apply_act(round_bf16(g), p.act)) * xrow[v * 8 + j];
            }
            if (vec_ok && n0 + v * 8 + 8 <= p.N) {
              uint4 o;
              o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]); o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
              *reinterpret_cast<uint4*>(crow + v * 8) = o;
            } else {
              for (int j = 0; j < 8 && n0 + v * 8 + j < p.N; ++j) crow[v * 8 + j] = __float2bfloat16_rn(f[j]);
            }
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");  // xchg is reused by the next tile
      } else if (row < p.M) {
        const int col0 = n0 + 64 * half;
        bf16* crow = p.C + (size_t)row * p.N + col0;
        const bf16* rrow = p.residual ? p.residual + (size_t)row * p.N + col0 : nullptr;
#pragma unroll
        for (int v = 0; v < 8; ++v) {
          if (col0 + v * 8 >= p.N) break;
          float f[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            f[j] = acc[v * 8 + j];
            if (p.bias && col0 + v * 8 + j < p.N) f[j] += __bfloat162float(p.bias[col0 + v * 8 + j]);
          }
          if (vec_ok && col0 + v * 8 + 8 <= p.N) {
            if (rrow) {  // eager semantics: linear output rounded to bf16, then added to the bf16 residual
              const uint4 r4 = *reinterpret_cast<const uint4*>(rrow + v * 8);
              const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                f[2 * j] = round_bf16(f[2 * j]) + bf16lo(rw[j]);
                f[2 * j + 1] = round_bf16(f[2 * j + 1]) + bf16hi(rw[j]);
              }
            }
            uint4 o;
            o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]); o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
            *reinterpret_cast<uint4*>(crow + v * 8) = o;
          } else {
            for (int j = 0; j < 8 && col0 + v * 8 + j < p.N; ++j) {
              float v0 = f[j];
              if (rrow) v0 = round_bf16(v0) + __bfloat162float(rrow[v * 8 + j]);
              crow[v * 8 + j] = __float2bfloat16_rn(v0);
            }
          }
        }
      }
    }
  }
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(F8_PBUF * F8_BN) : "memory");
  }
  hop_signal(p.signal, p.ctx);
}

// ---- activation quantiser: bf16 rows -> e4m3 + one fp32 scale per 128 K elements (stored transposed) -----------
// One warp per (row, group of K blocks): 4 elements per lane per block, amax by shuffle, scale = amax / 448.
__global__ void __launch_bounds__(256) quantize_rows_fp8_kernel(const bf16* __restrict__ x, unsigned char* __restrict__ q,
                                                                float* __restrict__ scale_t, int M, int K, long long ld_s) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x;
  const int nkb = K / 128;
  const bf16* xr = x + (size_t)row * K;
  unsigned char* qr = q + (size_t)row * K;
  for (int kb = warp; kb < nkb; kb += 8) {
    const uint2 raw = *reinterpret_cast<const uint2*>(xr + kb * 128 + lane * 4);
    const float v[4] = {bf16lo(raw.x), bf16hi(raw.x), bf16lo(raw.y), bf16hi(raw.y)};
    float amax = fmaxf(fmaxf(fabsf(v[0]), fabsf(v[1])), fmaxf(fabsf(v[2]), fabsf(v[3])));
    amax = warp_max(amax);
    const float scale = amax > 0.f ? amax * (1.f / 448.f) : 1.f;
    const float inv = 1.f / scale;
    uint32_t packed = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const __nv_fp8_storage_t b = __nv_cvt_float_to_fp8(v[j] * inv, __NV_SATFINITE, __NV_E4M3);
      packed |= (uint32_t)b << (8 * j);
    }
    *reinterpret_cast<uint32_t*>(qr + kb * 128 + lane * 4) = packed;
    if (lane == 0) scale_t[(size_t)kb * ld_s + row] = scale;
  }
}

// 2-D uint8 row-major [rows, cols] tensor, box = [box_rows, 128 bytes], 128-byte swizzle
static inline int make_map_u8(CUtensorMap* map, const void* ptr, long long rows, int cols, int box_rows) {
  PFN_encodeTiled enc = get_encode_fn();
  if (!enc) return -5;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols};
  cuuint32_t box[2] = {128u, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -6;
}

}  // namespace mdi

using namespace mdi;

// C[M,N] (bf16) = blockscaled(A8, a_scale_t) x blockscaled(W8, w_scale_t)^T (+bias)(+residual); with W2: gated MLP.
// K must be a multiple of 128.  signal_flag != null: fused prefill hop (epilogue stores go to peer memory).
extern "C" int mdi_gemm_fp8(const void* A8, const float* a_scale_t, long long ld_as, const void* W8, const float* w_scale_t,
                            const void* W2_8, const float* w2_scale_t, void* C, const void* bias, const void* bias2,
                            const void* residual, int M, int N, int K, int act, int* signal_flag, unsigned int* done_ctr,
                            const int* ctx, const int* status, cudaStream_t stream) {
  if (K % 128 != 0 || M <= 0 || N <= 0) return -2;
  const bool gated = W2_8 != nullptr;
  if (gated && residual) return -2;
  if (signal_flag && (!done_ctr || !ctx)) return -2;
  Fp8GemmParams p;
  p.a_scale_t = a_scale_t; p.w_scale_t = w_scale_t; p.w2_scale_t = w2_scale_t; p.ld_as = ld_as;
  p.C = (bf16*)C; p.bias = (const bf16*)bias; p.bias2 = (const bf16*)bias2; p.residual = (const bf16*)residual;
  p.M = M; p.N = N; p.K = K; p.act = act;
  p.signal = HopSignal{signal_flag, done_ctr, status}; p.ctx = ctx;
  int rc = make_map_u8(&p.tma_a, A8, M, K, F8_BM);
  if (rc) return rc;
  rc = make_map_u8(&p.tma_b, W8, N, K, gated ? 64 : F8_BN);
  if (rc) return rc;
  rc = make_map_u8(&p.tma_b2, gated ? W2_8 : W8, N, K, gated ? 64 : F8_BN);
  if (rc) return rc;
  if (N % 4 != 0) return -2;  // scale rows are fetched with 16-byte bulk copies
  const size_t smem = 1024 + (size_t)F8_STAGES * 2 * F8_TILE_BYTES + (size_t)F8_SCALE_SLOTS * 2 * 512 +
                      (gated ? (size_t)F8_BM * 64 * 4 : 0) + 256;
  const int n_cols = gated ? 64 : F8_BN;
  const int n_tiles = ((M + F8_BM - 1) / F8_BM) * ((N + n_cols - 1) / n_cols);
  dim3 grid(min(n_tiles, device_sm_count()));
  cudaError_t e;
  if (gated) {
    e = cudaFuncSetAttribute(gemm_fp8_blockscaled_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    gemm_fp8_blockscaled_kernel<true><<<grid, F8_THREADS, smem, stream>>>(p);
  } else {
    e = cudaFuncSetAttribute(gemm_fp8_blockscaled_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    gemm_fp8_blockscaled_kernel<false><<<grid, F8_THREADS, smem, stream>>>(p);
  }
  return (int)cudaGetLastError();
}

// x [M, K] bf16 -> q [M, K] e4m3 + scale_t [K/128, ld_s] fp32 (ld_s >= M)
extern "C" int mdi_quantize_rows_fp8(const void* x, void* q, float* scale_t, int M, int K, long long ld_s, cudaStream_t stream) {
  if (K % 128 != 0 || M <= 0 || ld_s < M) return -2;
  quantize_rows_fp8_kernel<<<M, 256, 0, stream>>>((const bf16*)x, (unsigned char*)q, scale_t, M, K, ld_s);
  return (int)cudaGetLastError();
}

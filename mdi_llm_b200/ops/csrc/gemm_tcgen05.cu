// bf16 GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), hand-written for sm_100a.
//
//   C[M, N] = A[M, K] · W[N, K]^T  (+ bias[N]) (+ residual[M, N])        bf16 in, fp32 accumulate
//
// This is the prefill path (T = prompt length rows through every linear of a stage — SURVEY K3/K8/
// K10/K11/K12 at T > 1), where the work is GEMM-shaped; decode (T = 1) is bandwidth-bound and uses
// the weight-streaming kernels in decode_linear.cu instead.
//
// Structure (one 128 x BLOCK_N output tile per CTA, warp-specialised, 4-stage TMA pipeline):
//   warp 0      : TMA producer — cp.async.bulk.tensor loads of the A and W tiles (64-element = 128 B
//                 K slabs, SWIZZLE_128B) into shared memory, completion on "full" mbarriers;
//   warp 1      : MMA issuer — one elected thread issues tcgen05.mma (M=128, N=BLOCK_N, K=16) from
//                 shared-memory descriptors into a TMEM accumulator; tcgen05.commit releases the
//                 stage ("empty" mbarrier) and finally signals the epilogue;
//   warp 2      : TMEM allocator / deallocator;
//   warps 4..7  : epilogue — tcgen05.ld (32 lanes x 32 columns per warp), bias/residual, bf16 store.
//
// Both operands are K-major (row-major with K contiguous), i.e. activations [T, C] and nn.Linear
// weights [out, in] are consumed as stored; out-of-range rows of the last M/N tile are zero-filled
// by TMA on load and masked on store.
#include "tcgen05.cuh"

namespace mdi {

constexpr int GEMM_BLOCK_M = 128;
constexpr int GEMM_BLOCK_K = 64;  // 64 bf16 = 128 bytes = one swizzle atom row
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_THREADS = 256;
constexpr int UMMA_K = 16;

struct GemmParams {
  CUtensorMap tma_a;
  CUtensorMap tma_b;
  CUtensorMap tma_b2;         // gated mode: second weight matrix (same shape as B)
  bf16* C;                     // may be peer-mapped memory: the epilogue's stores ARE the prefill hop
  const bf16* bias;
  const bf16* bias2;
  const bf16* residual;
  int M, N, K;
  int act;                     // gated mode (ACT_*): C = act(A W^T + bias) * (A W2^T + bias2)
  // fused hop: after its stores every CTA takes a ticket; the last one publishes flag[slot] = signal
  // (values from ctx, like the decode kernels) with a system-scope release.
  HopSignal signal;
  const int* ctx;
  // descriptor knobs (kept as parameters so a bring-up test can sweep them; defaults are the
  // canonical K-major SWIZZLE_128B encoding)
  unsigned int desc_sbo;      // stride-byte-offset >> 4 (8 rows x 128 B = 1024 B -> 64)
  unsigned int desc_lbo;      // leading-byte-offset >> 4 (ignored for swizzled K-major; 1)
  unsigned int desc_hi_bits;  // bits [46..63] >> 32 shifted: version (bit 46) | layout type (bits 61-63)
  unsigned int k_step_bytes;  // start-address advance per UMMA_K (16 bf16 = 32 B)
};

// shared-memory matrix descriptor: K-major, SWIZZLE_128B (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, const GemmParams& p) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);            // start address, bits [0,14)
  d |= (uint64_t)(p.desc_lbo & 0x3FFF) << 16;             // leading byte offset, bits [16,30)
  d |= (uint64_t)(p.desc_sbo & 0x3FFF) << 32;             // stride byte offset, bits [32,46)
  d |= (uint64_t)p.desc_hi_bits << 46;                    // version (bit 46) ... layout type (bits 61-63)
  return d;
}

// GATED: two B operands share the A tile; accumulators live side by side in TMEM (columns [0,BLOCK_N)
// and [BLOCK_N, 2*BLOCK_N)) and the epilogue writes act(acc0) * acc1 — SwiGLU / GeGLU in one pass.
//
// Persistent: one CTA per SM walks the tile list (m fastest, so the CTAs that run together share B tiles
// through L2).  The accumulator is double-buffered in TMEM: while the epilogue warps drain tile i
// (tcgen05.ld -> bias/act/residual -> 16-byte stores), the MMA warp is already accumulating tile i+1 in
// the other buffer and the TMA warp keeps the shared-memory ring full across tile boundaries.
template <int BLOCK_N, bool GATED>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_bf16_tcgen05_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ __align__(1024) unsigned char gemm_smem[];
  constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;
  constexpr int B_BYTES = (GATED ? 2 : 1) * BLOCK_N * GEMM_BLOCK_K * 2;
  constexpr int ACC_COLS = (GATED ? 2 : 1) * BLOCK_N;  // TMEM columns of one accumulator buffer
  constexpr int TMEM_COLS = 2 * ACC_COLS;              // two buffers (power of two, <= 512)
  static_assert(TMEM_COLS <= 512, "accumulators exceed TMEM");
  // carve: [A stages][B stages][barriers][tmem ptr]
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gemm_smem) + 1023) & ~uintptr_t(1023));
  unsigned char* smem_a = base;
  unsigned char* smem_b = base + GEMM_STAGES * A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + GEMM_STAGES * B_BYTES);
  uint64_t* empty_bar = full_bar + GEMM_STAGES;
  uint64_t* tmem_full_bar = empty_bar + GEMM_STAGES;  // [2]
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;        // [2]
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tiles_m = (p.M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M;
  const int tiles_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int n_tiles = tiles_m * tiles_n;
  const int num_k_blocks = (p.K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tma_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tma_b) : "memory");
    if (GATED) asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tma_b2) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < GEMM_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], 4); }
    mbar_fence_init();
  }
  if (warp == 2) {  // whole warp: allocate both accumulator buffers
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===== TMA producer: the ring runs straight through tile boundaries =====
    if (lane == 0) {
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int m0 = (tile % tiles_m) * GEMM_BLOCK_M, n0 = (tile / tiles_m) * BLOCK_N;
        for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
          const int s = it % GEMM_STAGES;
          const uint32_t phase = (it / GEMM_STAGES) & 1;
          mbar_wait(&empty_bar[s], phase ^ 1);
          mbar_expect_tx(&full_bar[s], A_BYTES + B_BYTES);
          tma_load_2d(smem_a + s * A_BYTES, &p.tma_a, &full_bar[s], kb * GEMM_BLOCK_K, m0);
          tma_load_2d(smem_b + s * B_BYTES, &p.tma_b, &full_bar[s], kb * GEMM_BLOCK_K, n0);
          if (GATED) tma_load_2d(smem_b + s * B_BYTES + B_BYTES / 2, &p.tma_b2, &full_bar[s], kb * GEMM_BLOCK_K, n0);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // instruction descriptor (cute::UMMA::InstrDescriptor): fp32 accumulate, bf16 x bf16, both K-major
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(GEMM_BLOCK_M >> 4) << 24);
    int it = 0, lt = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
      const int acc = lt & 1;
      const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * ACC_COLS);
      mbar_wait(&tmem_empty_bar[acc], ((lt >> 1) & 1) ^ 1);  // the epilogue has drained this buffer
      tcgen05_fence_after();
      for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
        const int s = it % GEMM_STAGES;
        const uint32_t phase = (it / GEMM_STAGES) & 1;
        mbar_wait(&full_bar[s], phase);
        tcgen05_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(smem_a + s * A_BYTES), b_addr = smem_u32(smem_b + s * B_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / UMMA_K; ++k) {
            const uint64_t da = make_smem_desc(a_addr + k * p.k_step_bytes, p);
            const uint64_t db = make_smem_desc(b_addr + k * p.k_step_bytes, p);
            tcgen05_mma_f16(tmem_acc, da, db, idesc, (kb | k) ? 1u : 0u);
            if (GATED) {
              const uint64_t db2 = make_smem_desc(b_addr + B_BYTES / 2 + k * p.k_step_bytes, p);
              tcgen05_mma_f16(tmem_acc + BLOCK_N, da, db2, idesc, (kb | k) ? 1u : 0u);
            }
          }
          tcgen05_commit(&empty_bar[s]);                            // stage reusable once these MMAs retire
          if (kb == num_k_blocks - 1) tcgen05_commit(&tmem_full_bar[acc]);  // accumulator complete
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: TMEM -> registers -> bf16 global =====
    const int ew = warp - 4;  // TMEM lanes [32*ew, 32*ew + 32) belong to this warp
    const bool vec_ok = (p.N % 8 == 0);
    int lt = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++lt) {
      const int m0 = (tile % tiles_m) * GEMM_BLOCK_M, n0 = (tile / tiles_m) * BLOCK_N;
      const int acc = lt & 1;
      const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * ACC_COLS) + ((uint32_t)(ew * 32) << 16);
      mbar_wait(&tmem_full_bar[acc], (lt >> 1) & 1);
      tcgen05_fence_after();
      const int row = m0 + ew * 32 + lane;
#pragma unroll 1
      for (int c0 = 0; c0 < BLOCK_N; c0 += 32) {
        uint32_t acc_r[32];
        tmem_ld_32x32(tmem_acc + (uint32_t)c0, acc_r);
        const int col0 = n0 + c0;
        if (GATED) {
          uint32_t acc2[32];
          tmem_ld_32x32(tmem_acc + (uint32_t)(BLOCK_N + c0), acc2);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float g = __uint_as_float(acc_r[j]), u = __uint_as_float(acc2[j]);
            if (col0 + j < p.N) {
              if (p.bias) g += __bfloat162float(p.bias[col0 + j]);
              if (p.bias2) u += __bfloat162float(p.bias2[col0 + j]);
            }
            // eager semantics: both projections are rounded to bf16 before the activation and the product
            g = round_bf16(g); u = round_bf16(u);
            acc_r[j] = __float_as_uint(round_bf16(apply_act(g, p.act)) * u);
          }
        }
        if (row < p.M && col0 < p.N) {
          bf16* crow = p.C + (size_t)row * p.N + col0;
          const bf16* rrow = p.residual ? p.residual + (size_t)row * p.N + col0 : nullptr;
          if (vec_ok && col0 + 32 <= p.N) {
            // 4 x 16-byte stores per thread (64 contiguous bytes of one output row): full sectors on the
            // wire when C is peer memory
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                f[j] = __uint_as_float(acc_r[v * 8 + j]);
                if (!GATED && p.bias) f[j] += __bfloat162float(p.bias[col0 + v * 8 + j]);
              }
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
              o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
              o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
              *reinterpret_cast<uint4*>(crow + v * 8) = o;
            }
          } else {
#pragma unroll 1
            for (int j = 0; j < 32 && col0 + j < p.N; ++j) {
              float v0 = __uint_as_float(acc_r[j]);
              if (!GATED && p.bias) v0 += __bfloat162float(p.bias[col0 + j]);
              if (rrow) v0 = round_bf16(v0) + __bfloat162float(rrow[j]);
              crow[j] = __float2bfloat16_rn(v0);
            }
          }
        }
      }
      // this warp is done reading the buffer: hand it back to the MMA warp (4 arrivals = 4 epilogue warps)
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty_bar[acc]);
    }
  }
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
  hop_signal(p.signal, p.ctx);  // the __syncthreads above ordered every epilogue store before the ticket
}

// ---- CTA-pair variant: tcgen05.mma.cta_group::2 -----------------------------------------------------------------
// Two CTAs of a cluster (same TPC) work on ONE 256 x BLOCK_N output tile: rank r owns rows [128 r, 128 r + 128) and loads
// its own A rows plus HALF of the B tile (BLOCK_N / 2 weight rows); the leader CTA issues M = 256 MMAs that read both
// CTAs' shared memory, every CTA's TMEM receives the accumulators of its own 128 rows.  Per CTA and K block that is
// 16 KB (A) + 16 KB (B half) for a 128 x 256 share of the output instead of 16 + 32 KB: a third less shared-memory
// fill and L2 traffic per FLOP, and the tensor core is fed M = 256.
//   * "full" barriers live in the leader; both CTAs' TMA loads complete on them (cp.async.bulk.tensor ... cta_group::2)
//   * "empty" / "accumulator full" barriers exist in both CTAs; the leader's tcgen05.commit multicasts to the pair
//   * "accumulator drained" lives in the leader; rank 1's epilogue warps arrive remotely (mapa + shared::cluster)
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* map, uint64_t* leader_bar, int x, int y) {
  // executed by both CTAs; clearing the peer bit makes the transaction bytes land on CTA 0's barrier
  const uint32_t bar = smem_u32(leader_bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(map), "r"(bar), "r"(x), "r"(y)
      : "memory");
}
__device__ __forceinline__ void tcgen05_mma_f16_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                                     uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit_pair(uint64_t* bar) {  // arrives on the same barrier in BOTH CTAs
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cta(uint64_t* bar, uint32_t cta) {  // arrive on `bar` of CTA `cta` of the cluster
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta) : "memory");
}

// GATED: the 256-column MMA tile is [fc_1 rows n0 .. n0+128 | fc_2 rows n0 .. n0+128] — rank 0 loads the fc_1 half of
// the B tile, rank 1 the fc_2 half — and the epilogue writes act(g) * u for 128 output columns.
constexpr int PAIR_STAGES = 6;  // 6 x (16 KB A + 16 KB B half) = 192 KB of the 227 KB
template <int BLOCK_N, bool GATED>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tcgen05_pair_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ __align__(1024) unsigned char gemm_smem[];
  constexpr int A_BYTES = GEMM_BLOCK_M * GEMM_BLOCK_K * 2;      // this CTA's 128 rows of A
  constexpr int B_BYTES = (BLOCK_N / 2) * GEMM_BLOCK_K * 2;     // this CTA's half of the B tile
  constexpr int TMEM_COLS = 2 * BLOCK_N;                        // two accumulator buffers
  static_assert(TMEM_COLS <= 512 && BLOCK_N % 32 == 0, "accumulators exceed TMEM");
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(gemm_smem) + 1023) & ~uintptr_t(1023));
  unsigned char* smem_a = base;
  unsigned char* smem_b = base + PAIR_STAGES * A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem_b + PAIR_STAGES * B_BYTES);  // used in the leader
  uint64_t* empty_bar = full_bar + PAIR_STAGES;                                       // both CTAs (multicast commit)
  uint64_t* tmem_full_bar = empty_bar + PAIR_STAGES;                                  // [2] both CTAs
  uint64_t* tmem_empty_bar = tmem_full_bar + 2;                                       // [2] leader: 8 epilogue warps of the pair
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  constexpr int OUT_N = GATED ? BLOCK_N / 2 : BLOCK_N;  // output columns per tile
  const int tiles_m = (p.M + 2 * GEMM_BLOCK_M - 1) / (2 * GEMM_BLOCK_M);
  const int tiles_n = (p.N + OUT_N - 1) / OUT_N;
  const int n_tiles = tiles_m * tiles_n;
  const int pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const int num_k_blocks = (p.K + GEMM_BLOCK_K - 1) / GEMM_BLOCK_K;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tma_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tma_b) : "memory");
    if (GATED) asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tma_b2) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < PAIR_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tmem_full_bar[b], 1); mbar_init(&tmem_empty_bar[b], 8); }
    mbar_fence_init();
  }
  if (warp == 2) {  // same warp in both CTAs: the allocation is collective over the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  cluster_sync_all();  // barriers of BOTH CTAs are initialised before any remote arrive / multicast commit
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own A rows + own half of B, completing on the leader's barrier =====
    if (lane == 0) {
      int it = 0;
      for (int tile = pair; tile < n_tiles; tile += n_pairs) {
        const int m0 = (tile % tiles_m) * 2 * GEMM_BLOCK_M + (int)rank * GEMM_BLOCK_M;
        const int n0 = GATED ? (tile / tiles_m) * OUT_N : (tile / tiles_m) * BLOCK_N + (int)rank * (BLOCK_N / 2);
        const CUtensorMap* map_b = (GATED && rank == 1) ? &p.tma_b2 : &p.tma_b;
        for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
          const int s = it % PAIR_STAGES;
          mbar_wait(&empty_bar[s], ((it / PAIR_STAGES) & 1) ^ 1);
          if (rank == 0) mbar_expect_tx(&full_bar[s], 2 * (A_BYTES + B_BYTES));  // bytes of the whole pair
          tma_load_2d_pair(smem_a + s * A_BYTES, &p.tma_a, &full_bar[s], kb * GEMM_BLOCK_K, m0);
          tma_load_2d_pair(smem_b + s * B_BYTES, map_b, &full_bar[s], kb * GEMM_BLOCK_K, n0);
        }
      }
    }
  } else if (warp == 1 && rank == 0) {
    // ===== MMA issuer (leader only): M = 256 over the pair =====
    const uint32_t idesc = umma_idesc_bf16(2 * GEMM_BLOCK_M, BLOCK_N);
    int it = 0, lt = 0;
    for (int tile = pair; tile < n_tiles; tile += n_pairs, ++lt) {
      const int acc = lt & 1;
      const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * BLOCK_N);
      mbar_wait(&tmem_empty_bar[acc], ((lt >> 1) & 1) ^ 1);  // both CTAs' epilogues have drained this buffer
      tcgen05_fence_after();
      for (int kb = 0; kb < num_k_blocks; ++kb, ++it) {
        const int s = it % PAIR_STAGES;
        mbar_wait(&full_bar[s], (it / PAIR_STAGES) & 1);
        tcgen05_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(smem_a + s * A_BYTES), b_addr = smem_u32(smem_b + s * B_BYTES);
#pragma unroll
          for (int k = 0; k < GEMM_BLOCK_K / UMMA_K; ++k)
            tcgen05_mma_f16_pair(tmem_acc, umma_desc_sw128(a_addr + k * 32), umma_desc_sw128(b_addr + k * 32), idesc, (kb | k) ? 1u : 0u);
          tcgen05_commit_pair(&empty_bar[s]);
          if (kb == num_k_blocks - 1) tcgen05_commit_pair(&tmem_full_bar[acc]);
        }
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue (both CTAs): own 128 rows =====
    const int ew = warp - 4;
    const bool vec_ok = (p.N % 8 == 0);
    int lt = 0;
    for (int tile = pair; tile < n_tiles; tile += n_pairs, ++lt) {
      const int m0 = (tile % tiles_m) * 2 * GEMM_BLOCK_M + (int)rank * GEMM_BLOCK_M, n0 = (tile / tiles_m) * OUT_N;
      const int acc = lt & 1;
      const uint32_t tmem_acc = tmem_base + (uint32_t)(acc * BLOCK_N) + ((uint32_t)(ew * 32) << 16);
      mbar_wait(&tmem_full_bar[acc], (lt >> 1) & 1);
      tcgen05_fence_after();
      const int row = m0 + ew * 32 + lane;
#pragma unroll 1
      for (int c0 = 0; c0 < OUT_N; c0 += 32) {
        uint32_t acc_r[32];
        tmem_ld_32x32(tmem_acc + (uint32_t)c0, acc_r);
        const int col0 = n0 + c0;
        if (GATED) {
          uint32_t acc2[32];
          tmem_ld_32x32(tmem_acc + (uint32_t)(OUT_N + c0), acc2);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float g = __uint_as_float(acc_r[j]), u = __uint_as_float(acc2[j]);
            if (col0 + j < p.N) {
              if (p.bias) g += __bfloat162float(p.bias[col0 + j]);
              if (p.bias2) u += __bfloat162float(p.bias2[col0 + j]);
            }
            g = round_bf16(g); u = round_bf16(u);
            acc_r[j] = __float_as_uint(round_bf16(apply_act(g, p.act)) * u);
          }
        }
        if (row < p.M && col0 < p.N) {
          bf16* crow = p.C + (size_t)row * p.N + col0;
          const bf16* rrow = p.residual ? p.residual + (size_t)row * p.N + col0 : nullptr;
          if (vec_ok && col0 + 32 <= p.N) {
#pragma unroll
            for (int v = 0; v < 4; ++v) {
              float f[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                f[j] = __uint_as_float(acc_r[v * 8 + j]);
                if (!GATED && p.bias) f[j] += __bfloat162float(p.bias[col0 + v * 8 + j]);
              }
              if (rrow) {
                const uint4 r4 = *reinterpret_cast<const uint4*>(rrow + v * 8);
                const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  f[2 * j] = round_bf16(f[2 * j]) + bf16lo(rw[j]);
                  f[2 * j + 1] = round_bf16(f[2 * j + 1]) + bf16hi(rw[j]);
                }
              }
              uint4 o;
              o.x = pack_bf16x2(f[0], f[1]); o.y = pack_bf16x2(f[2], f[3]);
              o.z = pack_bf16x2(f[4], f[5]); o.w = pack_bf16x2(f[6], f[7]);
              *reinterpret_cast<uint4*>(crow + v * 8) = o;
            }
          } else {
#pragma unroll 1
            for (int j = 0; j < 32 && col0 + j < p.N; ++j) {
              float v0 = __uint_as_float(acc_r[j]);
              if (!GATED && p.bias) v0 += __bfloat162float(p.bias[col0 + j]);
              if (rrow) v0 = round_bf16(v0) + __bfloat162float(rrow[j]);
              crow[j] = __float2bfloat16_rn(v0);
            }
          }
        }
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cta(&tmem_empty_bar[acc], 0);  // the leader's MMA warp waits for all 8 warps of the pair
    }
  }
  tcgen05_fence_before();
  cluster_sync_all();  // nobody leaves while the peer may still touch its barriers / shared memory
  if (warp == 2) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
  hop_signal(p.signal, p.ctx);
}

// ---- host side ------------------------------------------------------------------------------------
}  // namespace mdi

using namespace mdi;

// C[M,N] = A[M,K] W[N,K]^T (+bias) (+residual); with W2: C = act(A W^T + bias) * (A W2^T + bias2).
// K must be a multiple of 8 (16-byte rows for TMA).  signal_flag != null: fused hop (see GemmParams).
// knobs: sbo/lbo/hi_bits/k_step <= 0 select the canonical encoding.
extern "C" int mdi_gemm_bf16_ex(const void* A, const void* W, const void* W2, void* C, const void* bias,
                                const void* bias2, const void* residual, int M, int N, int K, int act, int block_n,
                                int* signal_flag, unsigned int* done_ctr, const int* ctx, const int* status, int sbo,
                                int lbo, int hi_bits, int k_step, cudaStream_t stream) {
  if (K % 8 != 0 || M <= 0 || N <= 0) return -2;
  if (W2 && residual) return -2;
  if (signal_flag && (!done_ctr || !ctx)) return -2;
  GemmParams p;
  p.C = (bf16*)C; p.bias = (const bf16*)bias; p.bias2 = (const bf16*)bias2; p.residual = (const bf16*)residual;
  p.M = M; p.N = N; p.K = K; p.act = act;
  p.signal = HopSignal{signal_flag, done_ctr, status}; p.ctx = ctx;
  p.desc_sbo = sbo > 0 ? (unsigned)sbo : 64u;
  p.desc_lbo = lbo > 0 ? (unsigned)lbo : 1u;
  p.desc_hi_bits = hi_bits > 0 ? (unsigned)hi_bits : (1u | (2u << 15));  // version = 1 (bit 46), SWIZZLE_128B = 2 (bits 61-63)
  p.k_step_bytes = k_step > 0 ? (unsigned)k_step : 32u;
  const bool gated = W2 != nullptr;
  if (block_n == 512) {  // CTA-pair kernel (cta_group::2): 256 x 256 MMA tiles shared by two CTAs of a cluster
    int rc2 = make_map(&p.tma_a, A, M, K, GEMM_BLOCK_M);
    if (rc2) return rc2;
    rc2 = make_map(&p.tma_b, W, N, K, 128);  // each CTA loads half (128 rows) of the 256-row B tile
    if (rc2) return rc2;
    rc2 = make_map(&p.tma_b2, gated ? W2 : W, N, K, 128);
    if (rc2) return rc2;
    const size_t smem2 = 1024 + (size_t)PAIR_STAGES * (GEMM_BLOCK_M + 128) * GEMM_BLOCK_K * 2 + 256;
    const int n_tiles2 = ((M + 255) / 256) * ((N + (gated ? 127 : 255)) / (gated ? 128 : 256));
    const int pairs = min(n_tiles2, device_sm_count() / 2);
    cudaError_t e2;
    if (gated) {
      e2 = cudaFuncSetAttribute(gemm_bf16_tcgen05_pair_kernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
      if (e2 != cudaSuccess) return (int)e2;
      gemm_bf16_tcgen05_pair_kernel<256, true><<<dim3(2 * pairs), GEMM_THREADS, smem2, stream>>>(p);
    } else {
      e2 = cudaFuncSetAttribute(gemm_bf16_tcgen05_pair_kernel<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
      if (e2 != cudaSuccess) return (int)e2;
      gemm_bf16_tcgen05_pair_kernel<256, false><<<dim3(2 * pairs), GEMM_THREADS, smem2, stream>>>(p);
    }
    return (int)cudaGetLastError();
  }
  if (block_n != 64 && block_n != 128 && block_n != 256) block_n = 128;
  if (gated && block_n == 256) block_n = 128;  // 2 buffers x 2 accumulators x BLOCK_N <= 512 TMEM columns
  int rc = make_map(&p.tma_a, A, M, K, GEMM_BLOCK_M);
  if (rc) return rc;
  rc = make_map(&p.tma_b, W, N, K, block_n);
  if (rc) return rc;
  rc = make_map(&p.tma_b2, gated ? W2 : W, N, K, block_n);
  if (rc) return rc;
  const size_t smem = 1024 + (size_t)GEMM_STAGES * (GEMM_BLOCK_M + (gated ? 2 : 1) * block_n) * GEMM_BLOCK_K * 2 + 128;
  const int n_tiles = ((M + GEMM_BLOCK_M - 1) / GEMM_BLOCK_M) * ((N + block_n - 1) / block_n);
  dim3 grid(min(n_tiles, device_sm_count()));  // persistent: one CTA per SM walks the tile list
  cudaError_t e;
#define MDI_GEMM_LAUNCH(BN, G)                                                                                        \
  e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BN, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
  if (e != cudaSuccess) return (int)e;                                                                               \
  gemm_bf16_tcgen05_kernel<BN, G><<<grid, GEMM_THREADS, smem, stream>>>(p);
  if (gated) { if (block_n == 64) { MDI_GEMM_LAUNCH(64, true) } else { MDI_GEMM_LAUNCH(128, true) } }
  else if (block_n == 64) { MDI_GEMM_LAUNCH(64, false) } else if (block_n == 256) { MDI_GEMM_LAUNCH(256, false) }
  else { MDI_GEMM_LAUNCH(128, false) }
#undef MDI_GEMM_LAUNCH
  return (int)cudaGetLastError();
}

extern "C" int mdi_gemm_bf16(const void* A, const void* W, void* C, const void* bias, const void* residual, int M,
                             int N, int K, int block_n, int sbo, int lbo, int hi_bits, int k_step,
                             cudaStream_t stream) {
  return mdi_gemm_bf16_ex(A, W, nullptr, C, bias, nullptr, residual, M, N, K, 0, block_n, nullptr, nullptr, nullptr,
                          nullptr, sbo, lbo, hi_bits, k_step, stream);
}

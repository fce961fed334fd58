// Causal flash attention for the prefill (T > 1) on the 5th-generation tensor cores, sm_100a.
//
// Reference behaviour being replaced (SURVEY K4-K7 at prefill): view/permute/split of the fused QKV
// output, RoPE as ~8 elementwise launches, index_copy_ into a cache expanded to n_head heads, then
// F.scaled_dot_product_attention with an explicit mask (model.py:693-751).  Here:
//
//   rope_split_kernel     one pass over the QKV GEMM's output [T, (H+2G)*hs]: RoPE on q and k, q written
//                         head-major for TMA, k/v appended to the slot KV pool (G group heads only) and v
//                         additionally transposed to [G, hs, T] so that P.V sees a K-major B operand;
//   attn_prefill_tcgen05  grid (T/128 query tiles, H heads).  Per 128-key tile:
//                         S = Q K^T      tcgen05.mma, operands staged by TMA (SWIZZLE_128B), S in TMEM
//                         softmax        4 warps, one query row per thread: tcgen05.ld, scale, causal
//                                        mask, online max/sum in registers, P (bf16) written to shared
//                                        memory in the same swizzled K-major layout TMA would produce
//                         O_t = P V      tcgen05.mma into a second TMEM region, added to the register
//                                        accumulator (already rescaled by exp2(m_old - m_new))
//                         one elected thread issues every TMA and MMA; mbarriers carry all hand-offs.
//
// GQA: the q heads of a group read the same K / V^T tiles (L2 hits after the first head).
#include "tcgen05.cuh"

namespace mdi {

constexpr int PA_BM = 128;       // query rows per CTA  (= TMEM lanes)
constexpr int PA_BN = 128;       // keys per tile
constexpr int PA_THREADS = 192;  // warps 0-3 softmax / epilogue, warp 4 TMA + MMA issue, warp 5 TMEM alloc

struct PrefillAttnParams {
  CUtensorMap map_q;   // [H * T_pad, hs]
  CUtensorMap map_k;   // this layer's KV pool viewed as [n_slots * 2 * G * S, hs]
  CUtensorMap map_vt;  // [G * hs, T_pad]
  bf16* y;             // [T, H * hs]
  int T, T_pad, n_head, q_per_kv, n_groups, max_seq, slot;
  float scale_log2;    // (1 / sqrt(hs)) * log2(e)
};

// ---- RoPE + split ------------------------------------------------------------------------------------
struct RopeSplitArgs {
  const bf16* qkv;   // [T, (H + 2G) * hs], litGPT group-interleaved columns
  const float* cos;  // [S, ne]
  const float* sin;
  bf16* q_out;       // [H, T_pad, hs]
  bf16* kv;          // layer pool [n_slots, 2, G, S, hs]
  bf16* vt;          // [G, hs, T_pad]
  int T, T_pad, slot, n_head, n_groups, hs, ne, max_seq;
};

// grid (ceil(T / 16), H + 2G): one head slot (a q head, the k head or the v head of a group) for 16 tokens.
// A thread owns 8 consecutive dims of one token: 16-byte loads / stores everywhere; V is additionally
// transposed through shared memory so that V^T rows are written 16 tokens (32 B) at a time.
constexpr int RS_TOK = 16;
__global__ void __launch_bounds__(256) rope_split_kernel(const RopeSplitArgs a) {
  __shared__ __align__(16) bf16 vt_tile[128][RS_TOK];
  const int hs = a.hs, ne = a.ne, half_ne = ne / 2;
  const int chunks = hs / 8;                       // 8-dim chunks per head (8 or 16)
  const int tt = threadIdx.x / chunks, ch = threadIdx.x % chunks;
  const int t = blockIdx.x * RS_TOK + tt, j = blockIdx.y;
  const int qpk = a.n_head / a.n_groups;
  const int g = j / (qpk + 2), sl = j % (qpk + 2);
  const bool active = tt < RS_TOK && t < a.T;
  const int d0 = ch * 8;
  uint4 out = make_uint4(0, 0, 0, 0);
  if (active) {
    const bf16* row = a.qkv + ((size_t)t * (a.n_head + 2 * a.n_groups) + j) * hs;
    const uint4 xv = *reinterpret_cast<const uint4*>(row + d0);
    out = xv;
    if (sl <= qpk && d0 < ne) {  // q and k heads: NeoX half rotation on the first `ne` dims (model.py:881-891)
      const bool lo = d0 < half_ne;
      const uint4 pv = *reinterpret_cast<const uint4*>(row + (lo ? d0 + half_ne : d0 - half_ne));
      const float4 c0 = *reinterpret_cast<const float4*>(a.cos + (size_t)t * ne + d0), c1 = *reinterpret_cast<const float4*>(a.cos + (size_t)t * ne + d0 + 4);
      const float4 s0 = *reinterpret_cast<const float4*>(a.sin + (size_t)t * ne + d0), s1 = *reinterpret_cast<const float4*>(a.sin + (size_t)t * ne + d0 + 4);
      const float cs[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
      const float sn[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
      const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w}, pw[4] = {pv.x, pv.y, pv.z, pv.w};
      uint32_t ow[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float x0 = bf16lo(xw[q]), x1 = bf16hi(xw[q]);
        const float r0 = lo ? -bf16lo(pw[q]) : bf16lo(pw[q]), r1 = lo ? -bf16hi(pw[q]) : bf16hi(pw[q]);
        ow[q] = pack_bf16x2(fmaf(x0, cs[2 * q], r0 * sn[2 * q]), fmaf(x1, cs[2 * q + 1], r1 * sn[2 * q + 1]));
      }
      out = make_uint4(ow[0], ow[1], ow[2], ow[3]);
    }
    if (sl < qpk) {
      *reinterpret_cast<uint4*>(a.q_out + ((size_t)(g * qpk + sl) * a.T_pad + t) * hs + d0) = out;
    } else {
      const size_t which = (sl == qpk) ? 0 : 1;
      *reinterpret_cast<uint4*>(a.kv + ((((size_t)a.slot * 2 + which) * a.n_groups + g) * a.max_seq + t) * hs + d0) = out;
    }
  }
  if (sl != qpk + 1) return;  // uniform per block: only V blocks transpose
  if (tt < RS_TOK) {
    const uint32_t w[4] = {out.x, out.y, out.z, out.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      vt_tile[d0 + 2 * q][tt] = __ushort_as_bfloat16((unsigned short)(w[q] & 0xffffu));
      vt_tile[d0 + 2 * q + 1][tt] = __ushort_as_bfloat16((unsigned short)(w[q] >> 16));
    }
  }
  __syncthreads();
  const int d = threadIdx.x / 2, part = threadIdx.x % 2;  // two threads per V^T row, 8 tokens (16 B) each
  if (d < hs)
    *reinterpret_cast<uint4*>(a.vt + ((size_t)g * hs + d) * a.T_pad + blockIdx.x * RS_TOK + part * 8) =
        *reinterpret_cast<const uint4*>(&vt_tile[d][part * 8]);
}

// ---- attention -----------------------------------------------------------------------------------------
template <int HS>
__global__ void __launch_bounds__(PA_THREADS, 1) attn_prefill_tcgen05_kernel(const __grid_constant__ PrefillAttnParams p) {
  extern __shared__ __align__(1024) unsigned char pa_smem[];
  constexpr int KA = HS / 64;                  // 64-column swizzle atoms along the head dimension
  constexpr int ROW_ATOM = 128 * 128;          // bytes of one [128 rows x 128 B] atom
  constexpr int Q_BYTES = KA * ROW_ATOM;
  constexpr int K_BYTES = KA * ROW_ATOM;
  constexpr int VT_ATOM = HS * 128;            // [HS rows x 64 keys]
  constexpr int VT_BYTES = 2 * VT_ATOM;
  constexpr int P_BYTES = 2 * ROW_ATOM;
  constexpr int TMEM_COLS = 256;               // S: columns [0,128), O_t: [128, 128 + HS)
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(pa_smem) + 1023) & ~uintptr_t(1023));
  unsigned char* q_s = base;
  unsigned char* k_s = q_s + Q_BYTES;
  unsigned char* vt_s = k_s + K_BYTES;
  unsigned char* p_s = vt_s + VT_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_s + P_BYTES);
  uint64_t *q_full = bars, *kv_full = bars + 1, *s_full = bars + 2, *p_ready = bars + 3, *o_full = bars + 4, *o_done = bars + 5;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 6);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x, h = blockIdx.y, g = h / p.q_per_kv;
  const int m0 = m_tile * PA_BM;
  const int n_tiles = m_tile + 1;  // causal: keys [0, m0 + 128)

  if (warp == 4 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.map_vt) : "memory");
    mbar_init(q_full, 1); mbar_init(kv_full, 1); mbar_init(s_full, 1); mbar_init(o_full, 1);
    mbar_init(p_ready, 128); mbar_init(o_done, 128);
    mbar_fence_init();
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_s = *tmem_ptr_smem, tmem_o = tmem_s + 128;

  if (warp == 4) {
    if (lane == 0) {
      // ===== TMA + MMA issue =====
      const uint32_t idesc_s = umma_idesc_bf16(PA_BM, PA_BN), idesc_o = umma_idesc_bf16(PA_BM, HS);
      const int q_row0 = h * p.T_pad + m0;
      const int k_row0 = ((p.slot * 2 + 0) * p.n_groups + g) * p.max_seq;
      mbar_expect_tx(q_full, Q_BYTES);
#pragma unroll
      for (int a = 0; a < KA; ++a) tma_load_2d(q_s + a * ROW_ATOM, &p.map_q, q_full, a * 64, q_row0);
      for (int t = 0; t < n_tiles; ++t) {
        const int j0 = t * PA_BN;
        if (t > 0) mbar_wait(o_full, (t - 1) & 1);  // previous P.V retired: K, V^T and P buffers are free
        mbar_expect_tx(kv_full, K_BYTES + VT_BYTES);
#pragma unroll
        for (int a = 0; a < KA; ++a) tma_load_2d(k_s + a * ROW_ATOM, &p.map_k, kv_full, a * 64, k_row0 + j0);
#pragma unroll
        for (int a = 0; a < 2; ++a) tma_load_2d(vt_s + a * VT_ATOM, &p.map_vt, kv_full, j0 + a * 64, g * HS);
        if (t == 0) mbar_wait(q_full, 0);
        mbar_wait(kv_full, t & 1);
        tcgen05_fence_after();
        // S = Q K^T : K dimension = head size
#pragma unroll
        for (int k = 0; k < HS / 16; ++k) {
          const uint32_t off = (uint32_t)(k / 4) * ROW_ATOM + (uint32_t)(k % 4) * 32;
          tcgen05_mma_f16(tmem_s, umma_desc_sw128(smem_u32(q_s) + off), umma_desc_sw128(smem_u32(k_s) + off), idesc_s, k ? 1u : 0u);
        }
        tcgen05_commit(s_full);
        mbar_wait(p_ready, t & 1);                  // P is in shared memory, S has been read
        if (t > 0) mbar_wait(o_done, (t - 1) & 1);  // previous O_t has been read out of TMEM
        tcgen05_fence_after();
        // O_t = P V : K dimension = the 128 keys of this tile
#pragma unroll
        for (int k = 0; k < PA_BN / 16; ++k) {
          const uint32_t offp = (uint32_t)(k / 4) * ROW_ATOM + (uint32_t)(k % 4) * 32;
          const uint32_t offv = (uint32_t)(k / 4) * VT_ATOM + (uint32_t)(k % 4) * 32;
          tcgen05_mma_f16(tmem_o, umma_desc_sw128(smem_u32(p_s) + offp), umma_desc_sw128(smem_u32(vt_s) + offv), idesc_o, k ? 1u : 0u);
        }
        tcgen05_commit(o_full);
      }
    }
  } else if (warp < 4) {
    // ===== softmax + output: thread <-> query row <-> TMEM lane =====
    const int r = threadIdx.x;
    const int q_idx = m0 + r;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    float O[HS];
#pragma unroll
    for (int d = 0; d < HS; ++d) O[d] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    for (int t = 0; t < n_tiles; ++t) {
      const int j0 = t * PA_BN;
      mbar_wait(s_full, t & 1);
      tcgen05_fence_after();
      // pass 1: row maximum of the masked, scaled scores
      float mx = -INFINITY;
#pragma unroll 1
      for (int c = 0; c < PA_BN / 32; ++c) {
        uint32_t sv[32];
        tmem_ld_32x32(tmem_s + lane_off + (uint32_t)(c * 32), sv);
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int j = j0 + c * 32 + i;
          if (j <= q_idx && j < p.T) mx = fmaxf(mx, __uint_as_float(sv[i]) * p.scale_log2);
        }
      }
      const float m_new = fmaxf(m_run, mx);
      const float corr = (m_run == -INFINITY) ? 0.f : exp2f(m_run - m_new);
      // pass 2: probabilities -> bf16 -> swizzled K-major tile in shared memory
      float psum = 0.f;
#pragma unroll 1
      for (int c = 0; c < PA_BN / 32; ++c) {
        uint32_t sv[32];
        tmem_ld_32x32(tmem_s + lane_off + (uint32_t)(c * 32), sv);
        float pf[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          const int j = j0 + c * 32 + i;
          const float pv = (j <= q_idx && j < p.T && m_new != -INFINITY) ? exp2f(__uint_as_float(sv[i]) * p.scale_log2 - m_new) : 0.f;
          pf[i] = pv;
          psum += pv;
        }
#pragma unroll
        for (int v = 0; v < 4; ++v) {  // 4 x 16 B = keys [c*32 + 8v, c*32 + 8v + 8)
          const int key = c * 32 + v * 8;
          const int atom = key >> 6, chunk = (key & 63) >> 3;
          uint4 o;
          o.x = pack_bf16x2(pf[v * 8 + 0], pf[v * 8 + 1]); o.y = pack_bf16x2(pf[v * 8 + 2], pf[v * 8 + 3]);
          o.z = pack_bf16x2(pf[v * 8 + 4], pf[v * 8 + 5]); o.w = pack_bf16x2(pf[v * 8 + 6], pf[v * 8 + 7]);
          *reinterpret_cast<uint4*>(p_s + atom * ROW_ATOM + r * 128 + ((chunk ^ (r & 7)) << 4)) = o;
        }
      }
      l_run = l_run * corr + psum;
      m_run = m_new;
#pragma unroll
      for (int d = 0; d < HS; ++d) O[d] *= corr;
      fence_proxy_async_smem();  // P stores (generic proxy) -> visible to the MMA's operand reads (async proxy)
      tcgen05_fence_before();
      mbar_arrive(p_ready);
      mbar_wait(o_full, t & 1);
      tcgen05_fence_after();
#pragma unroll
      for (int c = 0; c < HS / 32; ++c) {
        uint32_t ov[32];
        tmem_ld_32x32(tmem_o + lane_off + (uint32_t)(c * 32), ov);
#pragma unroll
        for (int i = 0; i < 32; ++i) O[c * 32 + i] += __uint_as_float(ov[i]);
      }
      tcgen05_fence_before();
      mbar_arrive(o_done);
    }
    if (q_idx < p.T) {
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      bf16* dst = p.y + (size_t)q_idx * p.n_head * HS + (size_t)h * HS;
#pragma unroll
      for (int v = 0; v < HS / 8; ++v) {
        uint4 o;
        o.x = pack_bf16x2(O[v * 8 + 0] * inv, O[v * 8 + 1] * inv); o.y = pack_bf16x2(O[v * 8 + 2] * inv, O[v * 8 + 3] * inv);
        o.z = pack_bf16x2(O[v * 8 + 4] * inv, O[v * 8 + 5] * inv); o.w = pack_bf16x2(O[v * 8 + 6] * inv, O[v * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(dst + v * 8) = o;
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 5) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_s), "n"(TMEM_COLS) : "memory");
  }
}

// ---- pipelined variant ------------------------------------------------------------------------------------
// Same math, deeper overlap: K / V^T tiles are double-buffered in shared memory (the TMA of tile t+2 is in
// flight while tile t+1 is being used) and the score matrix is double-buffered in TMEM (S[t+1] = Q K^T is
// computed by the tensor cores WHILE the softmax warps work on S[t]).  Per tile only the softmax, the
// P.V MMA and the O_t read-out remain on the serial chain.
template <int HS>
__global__ void __launch_bounds__(PA_THREADS, 1) attn_prefill_tcgen05_pipe_kernel(const __grid_constant__ PrefillAttnParams p) {
  extern __shared__ __align__(1024) unsigned char pa_smem[];
  constexpr int KA = HS / 64;
  constexpr int ROW_ATOM = 128 * 128;
  constexpr int Q_BYTES = KA * ROW_ATOM;
  constexpr int K_BYTES = KA * ROW_ATOM;
  constexpr int VT_ATOM = HS * 128;
  constexpr int VT_BYTES = 2 * VT_ATOM;
  constexpr int P_BYTES = 2 * ROW_ATOM;
  constexpr int TMEM_COLS = 512;  // S0: [0,128)  S1: [128,256)  O_t: [256, 256 + HS)
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(pa_smem) + 1023) & ~uintptr_t(1023));
  unsigned char* q_s = base;
  unsigned char* k_s = q_s + Q_BYTES;            // [2][K_BYTES]
  unsigned char* vt_s = k_s + 2 * K_BYTES;       // [2][VT_BYTES]
  unsigned char* p_s = vt_s + 2 * VT_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_s + P_BYTES);
  uint64_t *q_full = bars, *kv_full = bars + 1 /* [2] */, *s_full = bars + 3 /* [2] */, *p_ready = bars + 5,
           *o_full = bars + 6, *o_done = bars + 7;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 8);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x, h = blockIdx.y, g = h / p.q_per_kv;
  const int m0 = m_tile * PA_BM;
  const int n_tiles = m_tile + 1;

  if (warp == 4 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.map_vt) : "memory");
    mbar_init(q_full, 1); mbar_init(o_full, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(&kv_full[b], 1); mbar_init(&s_full[b], 1); }
    mbar_init(p_ready, 128); mbar_init(o_done, 128);
    mbar_fence_init();
  }
  if (warp == 5) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem, tmem_o = tmem_base + 256;

  if (warp == 4) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(PA_BM, PA_BN), idesc_o = umma_idesc_bf16(PA_BM, HS);
      const int q_row0 = h * p.T_pad + m0;
      const int k_row0 = ((p.slot * 2 + 0) * p.n_groups + g) * p.max_seq;
      auto load_tile = [&](int t) {  // K and V^T of key tile t into buffer t & 1
        const int b = t & 1, j0 = t * PA_BN;
        mbar_expect_tx(&kv_full[b], K_BYTES + VT_BYTES);
#pragma unroll
        for (int a = 0; a < KA; ++a) tma_load_2d(k_s + b * K_BYTES + a * ROW_ATOM, &p.map_k, &kv_full[b], a * 64, k_row0 + j0);
#pragma unroll
        for (int a = 0; a < 2; ++a) tma_load_2d(vt_s + b * VT_BYTES + a * VT_ATOM, &p.map_vt, &kv_full[b], j0 + a * 64, g * HS);
      };
      auto mma_s = [&](int t) {  // S[t & 1] = Q K_t^T
        const int b = t & 1;
        mbar_wait(&kv_full[b], (t >> 1) & 1);
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < HS / 16; ++k) {
          const uint32_t off = (uint32_t)(k / 4) * ROW_ATOM + (uint32_t)(k % 4) * 32;
          tcgen05_mma_f16(tmem_base + (uint32_t)(b * 128), umma_desc_sw128(smem_u32(q_s) + off),
                          umma_desc_sw128(smem_u32(k_s + b * K_BYTES) + off), idesc_s, k ? 1u : 0u);
        }
        tcgen05_commit(&s_full[b]);
      };
      mbar_expect_tx(q_full, Q_BYTES);
#pragma unroll
      for (int a = 0; a < KA; ++a) tma_load_2d(q_s + a * ROW_ATOM, &p.map_q, q_full, a * 64, q_row0);
      load_tile(0);
      if (n_tiles > 1) load_tile(1);
      mbar_wait(q_full, 0);
      mma_s(0);
      for (int t = 0; t < n_tiles; ++t) {
        const int b = t & 1;
        // (A) scores of the NEXT tile while the softmax warps are busy with this one.  S[(t+1)&1] is free:
        //     its previous reader (softmax of tile t-1) arrived on p_ready(t-1), waited for in (B) below.
        if (t + 1 < n_tiles) mma_s(t + 1);
        // (B) O_t = P V_t
        mbar_wait(p_ready, t & 1);
        if (t > 0) mbar_wait(o_done, (t - 1) & 1);
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < PA_BN / 16; ++k) {
          const uint32_t offp = (uint32_t)(k / 4) * ROW_ATOM + (uint32_t)(k % 4) * 32;
          const uint32_t offv = (uint32_t)(k / 4) * VT_ATOM + (uint32_t)(k % 4) * 32;
          tcgen05_mma_f16(tmem_o, umma_desc_sw128(smem_u32(p_s) + offp), umma_desc_sw128(smem_u32(vt_s + b * VT_BYTES) + offv),
                          idesc_o, k ? 1u : 0u);
        }
        tcgen05_commit(o_full);
        // (C) buffer b is free once P V_t has retired (K_t was consumed by S[t] long ago): fetch tile t + 2
        if (t + 2 < n_tiles) {
          mbar_wait(o_full, t & 1);
          load_tile(t + 2);
        }
      }
    }
  } else if (warp < 4) {
    const int r = threadIdx.x;
    const int q_idx = m0 + r;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    float O[HS];
#pragma unroll
    for (int d = 0; d < HS; ++d) O[d] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    for (int t = 0; t < n_tiles; ++t) {
      const int j0 = t * PA_BN;
      const uint32_t tmem_s = tmem_base + (uint32_t)((t & 1) * 128) + lane_off;
      mbar_wait(&s_full[t & 1], (t >> 1) & 1);
      tcgen05_fence_after();
      // only the diagonal tile (and a ragged last one) needs the causal / length mask; four independent
      // max / sum chains keep the single warp per scheduler from serialising on FP latency
      const bool full = (j0 + PA_BN - 1 <= m0) && (j0 + PA_BN <= p.T);
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
      for (int c = 0; c < PA_BN / 32; ++c) {
        uint32_t sv[32];
        tmem_ld_32x32(tmem_s + (uint32_t)(c * 32), sv);
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(sv[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int j = j0 + c * 32 + i;
            if (j <= q_idx && j < p.T) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(sv[i]));
          }
        }
      }
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * p.scale_log2;  // scale > 0: max commutes
      const float m_new = fmaxf(m_run, mx);
      const float corr = (m_run == -INFINITY) ? 0.f : exp2f(m_run - m_new);
      float ps4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int c = 0; c < PA_BN / 32; ++c) {
        uint32_t sv[32];
        tmem_ld_32x32(tmem_s + (uint32_t)(c * 32), sv);
        float pf[32];
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            pf[i] = exp2f(fmaf(__uint_as_float(sv[i]), p.scale_log2, -m_new));
            ps4[i & 3] += pf[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int j = j0 + c * 32 + i;
            pf[i] = (j <= q_idx && j < p.T && m_new != -INFINITY) ? exp2f(fmaf(__uint_as_float(sv[i]), p.scale_log2, -m_new)) : 0.f;
            ps4[i & 3] += pf[i];
          }
        }
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int key = c * 32 + v * 8;
          const int atom = key >> 6, chunk = (key & 63) >> 3;
          uint4 o;
          o.x = pack_bf16x2(pf[v * 8 + 0], pf[v * 8 + 1]); o.y = pack_bf16x2(pf[v * 8 + 2], pf[v * 8 + 3]);
          o.z = pack_bf16x2(pf[v * 8 + 4], pf[v * 8 + 5]); o.w = pack_bf16x2(pf[v * 8 + 6], pf[v * 8 + 7]);
          *reinterpret_cast<uint4*>(p_s + atom * ROW_ATOM + r * 128 + ((chunk ^ (r & 7)) << 4)) = o;
        }
      }
      l_run = l_run * corr + ((ps4[0] + ps4[1]) + (ps4[2] + ps4[3]));
      m_run = m_new;
#pragma unroll
      for (int d = 0; d < HS; ++d) O[d] *= corr;
      fence_proxy_async_smem();
      tcgen05_fence_before();
      mbar_arrive(p_ready);
      mbar_wait(o_full, t & 1);  // also guarantees P is free before the next tile's softmax overwrites it
      tcgen05_fence_after();
#pragma unroll
      for (int c = 0; c < HS / 32; ++c) {
        uint32_t ov[32];
        tmem_ld_32x32(tmem_o + lane_off + (uint32_t)(c * 32), ov);
#pragma unroll
        for (int i = 0; i < 32; ++i) O[c * 32 + i] += __uint_as_float(ov[i]);
      }
      tcgen05_fence_before();
      mbar_arrive(o_done);
    }
    if (q_idx < p.T) {
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      bf16* dst = p.y + (size_t)q_idx * p.n_head * HS + (size_t)h * HS;
#pragma unroll
      for (int v = 0; v < HS / 8; ++v) {
        uint4 o;
        o.x = pack_bf16x2(O[v * 8 + 0] * inv, O[v * 8 + 1] * inv); o.y = pack_bf16x2(O[v * 8 + 2] * inv, O[v * 8 + 3] * inv);
        o.z = pack_bf16x2(O[v * 8 + 4] * inv, O[v * 8 + 5] * inv); o.w = pack_bf16x2(O[v * 8 + 6] * inv, O[v * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(dst + v * 8) = o;
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 5) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

// ---- pipelined variant with two softmax warpgroups (EXPERIMENTAL: compiled, not yet validated on hardware) ----
// The softmax is the bottleneck of the pipelined kernel (one query row per thread, 4 warps).  Here 8 warps share it:
// warpgroup w (warps 4w..4w+3) owns key columns [64w, 64w+64) of every score tile and output dims [HS/2*w, HS/2*(w+1));
// the two threads of a row exchange their partial row maximum through shared memory (one named barrier per tile),
// keep partial row sums (same rescaling on both sides, added once at the end) and each write their 64 keys of P —
// which is exactly one swizzle atom per warpgroup.  Selected with mdi_set_prefill_attn_pipe(2).
// Same math, deeper overlap: K / V^T tiles are double-buffered in shared memory (the TMA of tile t+2 is in
// flight while tile t+1 is being used) and the score matrix is double-buffered in TMEM (S[t+1] = Q K^T is
// computed by the tensor cores WHILE the softmax warps work on S[t]).  Per tile only the softmax, the
// P.V MMA and the O_t read-out remain on the serial chain.
constexpr int PA2_THREADS = 320;  // warps 0-7 softmax (two warpgroups), warp 8 TMA + MMA issue, warp 9 TMEM alloc
template <int HS>
__global__ void __launch_bounds__(PA2_THREADS, 1) attn_prefill_tcgen05_pipe2_kernel(const __grid_constant__ PrefillAttnParams p) {
  extern __shared__ __align__(1024) unsigned char pa_smem[];
  constexpr int KA = HS / 64;
  constexpr int ROW_ATOM = 128 * 128;
  constexpr int Q_BYTES = KA * ROW_ATOM;
  constexpr int K_BYTES = KA * ROW_ATOM;
  constexpr int VT_ATOM = HS * 128;
  constexpr int VT_BYTES = 2 * VT_ATOM;
  constexpr int P_BYTES = 2 * ROW_ATOM;
  constexpr int TMEM_COLS = 512;  // S0: [0,128)  S1: [128,256)  O_t: [256, 256 + HS)
  unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(pa_smem) + 1023) & ~uintptr_t(1023));
  unsigned char* q_s = base;
  unsigned char* k_s = q_s + Q_BYTES;            // [2][K_BYTES]
  unsigned char* vt_s = k_s + 2 * K_BYTES;       // [2][VT_BYTES]
  unsigned char* p_s = vt_s + 2 * VT_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(p_s + P_BYTES);
  uint64_t *q_full = bars, *kv_full = bars + 1 /* [2] */, *s_full = bars + 3 /* [2] */, *p_ready = bars + 5,
           *o_full = bars + 6, *o_done = bars + 7;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(bars + 8);
  float* mx_x = reinterpret_cast<float*>(bars + 9);  // [2 tiles parity][2 warpgroups][128 rows] partial row maxima
  float* l_x = mx_x + 2 * 2 * 128;                   // [2 warpgroups][128 rows] partial row sums (final exchange)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x, h = blockIdx.y, g = h / p.q_per_kv;
  const int m0 = m_tile * PA_BM;
  const int n_tiles = m_tile + 1;

  if (warp == 8 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.map_q) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.map_k) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.map_vt) : "memory");
    mbar_init(q_full, 1); mbar_init(o_full, 1);
    for (int b = 0; b < 2; ++b) { mbar_init(&kv_full[b], 1); mbar_init(&s_full[b], 1); }
    mbar_init(p_ready, 256); mbar_init(o_done, 256);
    mbar_fence_init();
  }
  if (warp == 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem, tmem_o = tmem_base + 256;

  if (warp == 8) {
    if (lane == 0) {
      const uint32_t idesc_s = umma_idesc_bf16(PA_BM, PA_BN), idesc_o = umma_idesc_bf16(PA_BM, HS);
      const int q_row0 = h * p.T_pad + m0;
      const int k_row0 = ((p.slot * 2 + 0) * p.n_groups + g) * p.max_seq;
      auto load_tile = [&](int t) {  // K and V^T of key tile t into buffer t & 1
        const int b = t & 1, j0 = t * PA_BN;
        mbar_expect_tx(&kv_full[b], K_BYTES + VT_BYTES);
#pragma unroll
        for (int a = 0; a < KA; ++a) tma_load_2d(k_s + b * K_BYTES + a * ROW_ATOM, &p.map_k, &kv_full[b], a * 64, k_row0 + j0);
#pragma unroll
        for (int a = 0; a < 2; ++a) tma_load_2d(vt_s + b * VT_BYTES + a * VT_ATOM, &p.map_vt, &kv_full[b], j0 + a * 64, g * HS);
      };
      auto mma_s = [&](int t) {  // S[t & 1] = Q K_t^T
        const int b = t & 1;
        mbar_wait(&kv_full[b], (t >> 1) & 1);
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < HS / 16; ++k) {
          const uint32_t off = (uint32_t)(k / 4) * ROW_ATOM + (uint32_t)(k % 4) * 32;
          tcgen05_mma_f16(tmem_base + (uint32_t)(b * 128), umma_desc_sw128(smem_u32(q_s) + off),
                          umma_desc_sw128(smem_u32(k_s + b * K_BYTES) + off), idesc_s, k ? 1u : 0u);
        }
        tcgen05_commit(&s_full[b]);
      };
      mbar_expect_tx(q_full, Q_BYTES);
#pragma unroll
      for (int a = 0; a < KA; ++a) tma_load_2d(q_s + a * ROW_ATOM, &p.map_q, q_full, a * 64, q_row0);
      load_tile(0);
      if (n_tiles > 1) load_tile(1);
      mbar_wait(q_full, 0);
      mma_s(0);
      for (int t = 0; t < n_tiles; ++t) {
        const int b = t & 1;
        // (A) scores of the NEXT tile while the softmax warps are busy with this one.  S[(t+1)&1] is free:
        //     its previous reader (softmax of tile t-1) arrived on p_ready(t-1), waited for in (B) below.
        if (t + 1 < n_tiles) mma_s(t + 1);
        // (B) O_t = P V_t
        mbar_wait(p_ready, t & 1);
        if (t > 0) mbar_wait(o_done, (t - 1) & 1);
        tcgen05_fence_after();
#pragma unroll
        for (int k = 0; k < PA_BN / 16; ++k) {
          const uint32_t offp = (uint32_t)(k / 4) * ROW_ATOM + (uint32_t)(k % 4) * 32;
          const uint32_t offv = (uint32_t)(k / 4) * VT_ATOM + (uint32_t)(k % 4) * 32;
          tcgen05_mma_f16(tmem_o, umma_desc_sw128(smem_u32(p_s) + offp), umma_desc_sw128(smem_u32(vt_s + b * VT_BYTES) + offv),
                          idesc_o, k ? 1u : 0u);
        }
        tcgen05_commit(o_full);
        // (C) buffer b is free once P V_t has retired (K_t was consumed by S[t] long ago): fetch tile t + 2
        if (t + 2 < n_tiles) {
          mbar_wait(o_full, t & 1);
          load_tile(t + 2);
        }
      }
    }
  } else if (warp < 8) {
    constexpr int HD = HS / 2;                 // output dims per thread
    const int wg = warp >> 2;                  // warpgroup: key columns [64 wg, 64 wg + 64), output dims [HD wg, HD wg + HD)
    const int r = threadIdx.x & 127;           // query row == TMEM lane
    const int q_idx = m0 + r;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    const int cb = wg * 64;                    // first key column of this thread inside a tile
    float O[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) O[d] = 0.f;
    float m_run = -INFINITY, l_run = 0.f;
    for (int t = 0; t < n_tiles; ++t) {
      const int j0 = t * PA_BN;
      const uint32_t tmem_s = tmem_base + (uint32_t)((t & 1) * 128) + lane_off;
      mbar_wait(&s_full[t & 1], (t >> 1) & 1);
      tcgen05_fence_after();
      // only the diagonal tile (and a ragged last one) needs the causal / length mask; four independent
      // max / sum chains keep the single warp per scheduler from serialising on FP latency
      const bool full = (j0 + PA_BN - 1 <= m0) && (j0 + PA_BN <= p.T);
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t sv[32];
        tmem_ld_32x32(tmem_s + (uint32_t)(cb + c * 32), sv);
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(sv[i]));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int j = j0 + cb + c * 32 + i;
            if (j <= q_idx && j < p.T) mx4[i & 3] = fmaxf(mx4[i & 3], __uint_as_float(sv[i]));
          }
        }
      }
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * p.scale_log2;  // scale > 0: max commutes
      // the row maximum is over all 128 keys: exchange the two halves (parity-indexed slot, one named barrier)
      float* slot = mx_x + (t & 1) * 256;
      slot[wg * 128 + r] = mx;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mx = fmaxf(mx, slot[(wg ^ 1) * 128 + r]);
      const float m_new = fmaxf(m_run, mx);
      const float corr = (m_run == -INFINITY) ? 0.f : exp2f(m_run - m_new);
      float ps4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t sv[32];
        tmem_ld_32x32(tmem_s + (uint32_t)(cb + c * 32), sv);
        float pf[32];
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            pf[i] = exp2f(fmaf(__uint_as_float(sv[i]), p.scale_log2, -m_new));
            ps4[i & 3] += pf[i];
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int j = j0 + cb + c * 32 + i;
            pf[i] = (j <= q_idx && j < p.T && m_new != -INFINITY) ? exp2f(fmaf(__uint_as_float(sv[i]), p.scale_log2, -m_new)) : 0.f;
            ps4[i & 3] += pf[i];
          }
        }
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const int key = cb + c * 32 + v * 8;
          const int atom = key >> 6, chunk = (key & 63) >> 3;
          uint4 o;
          o.x = pack_bf16x2(pf[v * 8 + 0], pf[v * 8 + 1]); o.y = pack_bf16x2(pf[v * 8 + 2], pf[v * 8 + 3]);
          o.z = pack_bf16x2(pf[v * 8 + 4], pf[v * 8 + 5]); o.w = pack_bf16x2(pf[v * 8 + 6], pf[v * 8 + 7]);
          *reinterpret_cast<uint4*>(p_s + atom * ROW_ATOM + r * 128 + ((chunk ^ (r & 7)) << 4)) = o;
        }
      }
      l_run = l_run * corr + ((ps4[0] + ps4[1]) + (ps4[2] + ps4[3]));
      m_run = m_new;
#pragma unroll
      for (int d = 0; d < HD; ++d) O[d] *= corr;
      fence_proxy_async_smem();
      tcgen05_fence_before();
      mbar_arrive(p_ready);
      mbar_wait(o_full, t & 1);  // also guarantees P is free before the next tile's softmax overwrites it
      tcgen05_fence_after();
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t ov[32];
        tmem_ld_32x32(tmem_o + lane_off + (uint32_t)(wg * HD + c * 32), ov);
#pragma unroll
        for (int i = 0; i < 32; ++i) O[c * 32 + i] += __uint_as_float(ov[i]);
      }
      tcgen05_fence_before();
      mbar_arrive(o_done);
    }
    // total row sum = the two partial sums (both were rescaled by the same factors all along)
    l_x[wg * 128 + r] = l_run;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    l_run += l_x[(wg ^ 1) * 128 + r];
    if (q_idx < p.T) {
      const float inv = l_run > 0.f ? 1.f / l_run : 0.f;
      bf16* dst = p.y + (size_t)q_idx * p.n_head * HS + (size_t)h * HS + wg * HD;
#pragma unroll
      for (int v = 0; v < HD / 8; ++v) {
        uint4 o;
        o.x = pack_bf16x2(O[v * 8 + 0] * inv, O[v * 8 + 1] * inv); o.y = pack_bf16x2(O[v * 8 + 2] * inv, O[v * 8 + 3] * inv);
        o.z = pack_bf16x2(O[v * 8 + 4] * inv, O[v * 8 + 5] * inv); o.w = pack_bf16x2(O[v * 8 + 6] * inv, O[v * 8 + 7] * inv);
        *reinterpret_cast<uint4*>(dst + v * 8) = o;
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 9) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
}

static int g_prefill_attn_pipe = 2;  // two softmax warpgroups (validated on B200); see mdi_set_prefill_attn_pipe

template <int HS>
static int launch_prefill_attn_pipe(const PrefillAttnParams& p, cudaStream_t stream) {
  const size_t smem = 1024 + (size_t)((HS / 64) * 128 * 128 * 3 + 2 * 2 * HS * 128 + 2 * 128 * 128) + 128;
  cudaError_t e = cudaFuncSetAttribute(attn_prefill_tcgen05_pipe_kernel<HS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((p.T + PA_BM - 1) / PA_BM, p.n_head);
  attn_prefill_tcgen05_pipe_kernel<HS><<<grid, PA_THREADS, smem, stream>>>(p);
  return (int)cudaGetLastError();
}

template <int HS>
static int launch_prefill_attn_pipe2(const PrefillAttnParams& p, cudaStream_t stream) {
  const size_t smem = 1024 + (size_t)((HS / 64) * 128 * 128 * 3 + 2 * 2 * HS * 128 + 2 * 128 * 128) + 128 + (4 + 2) * 128 * 4;
  cudaError_t e = cudaFuncSetAttribute(attn_prefill_tcgen05_pipe2_kernel<HS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((p.T + PA_BM - 1) / PA_BM, p.n_head);
  attn_prefill_tcgen05_pipe2_kernel<HS><<<grid, PA2_THREADS, smem, stream>>>(p);
  return (int)cudaGetLastError();
}

template <int HS>
static int launch_prefill_attn(const PrefillAttnParams& p, cudaStream_t stream) {
  const size_t smem = 1024 + (size_t)(2 * (HS / 64) * 128 * 128 + 2 * HS * 128 + 2 * 128 * 128) + 128;
  cudaError_t e = cudaFuncSetAttribute(attn_prefill_tcgen05_kernel<HS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  dim3 grid((p.T + PA_BM - 1) / PA_BM, p.n_head);
  attn_prefill_tcgen05_kernel<HS><<<grid, PA_THREADS, smem, stream>>>(p);
  return (int)cudaGetLastError();
}

}  // namespace mdi

using namespace mdi;

// qkv: [T, (H + 2G) * hs] bf16 (output of the QKV GEMM), positions 0..T-1.  q_scratch: [H, T_pad, hs],
// vt_scratch: [G, hs, T_pad] (T_pad = T rounded up to 128, both finite — zero-initialised once),
// kv: the layer's pool [n_slots, 2, G, S, hs], y: [T, H * hs].
extern "C" int mdi_attn_prefill(const void* qkv, const float* cos, const float* sin, void* kv, void* q_scratch,
                                void* vt_scratch, void* y, int T, int T_pad, int slot, int n_slots, int n_head,
                                int n_groups, int head_size, int rope_n_elem, int max_seq, cudaStream_t stream) {
  if (T <= 0 || T > max_seq || T_pad % PA_BN != 0 || T_pad < T || n_head % n_groups != 0) return -2;
  if (head_size != 64 && head_size != 128) return -3;
  if (rope_n_elem % 16 != 0 || rope_n_elem > head_size) return -2;  // the vectorised pre-pass rotates 8-dim chunks
  RopeSplitArgs r;
  r.qkv = (const bf16*)qkv; r.cos = cos; r.sin = sin; r.q_out = (bf16*)q_scratch; r.kv = (bf16*)kv; r.vt = (bf16*)vt_scratch;
  r.T = T; r.T_pad = T_pad; r.slot = slot; r.n_head = n_head; r.n_groups = n_groups; r.hs = head_size; r.ne = rope_n_elem;
  r.max_seq = max_seq;
  rope_split_kernel<<<dim3((T + RS_TOK - 1) / RS_TOK, n_head + 2 * n_groups), 256, 0, stream>>>(r);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return (int)e;

  PrefillAttnParams p;
  p.y = (bf16*)y; p.T = T; p.T_pad = T_pad; p.n_head = n_head; p.n_groups = n_groups; p.q_per_kv = n_head / n_groups;
  p.max_seq = max_seq; p.slot = slot;
  p.scale_log2 = 1.4426950408889634f / sqrtf((float)head_size);
  int rc = make_map(&p.map_q, q_scratch, (long long)n_head * T_pad, head_size, PA_BM);
  if (rc) return rc;
  rc = make_map(&p.map_k, kv, (long long)n_slots * 2 * n_groups * max_seq, head_size, PA_BN);
  if (rc) return rc;
  rc = make_map(&p.map_vt, vt_scratch, (long long)n_groups * head_size, T_pad, head_size);
  if (rc) return rc;
  if (g_prefill_attn_pipe == 2) return head_size == 128 ? launch_prefill_attn_pipe2<128>(p, stream) : launch_prefill_attn_pipe2<64>(p, stream);
  if (g_prefill_attn_pipe) return head_size == 128 ? launch_prefill_attn_pipe<128>(p, stream) : launch_prefill_attn_pipe<64>(p, stream);
  return head_size == 128 ? launch_prefill_attn<128>(p, stream) : launch_prefill_attn<64>(p, stream);
}

// 0 = simple sequential kernel, 1 = pipelined (double-buffered K/V^T tiles and score matrix; default),
// 2 = pipelined with two softmax warpgroups (experimental)
extern "C" void mdi_set_prefill_attn_pipe(int on) { g_prefill_attn_pipe = on; }
extern "C" int mdi_get_prefill_attn_pipe() { return g_prefill_attn_pipe; }

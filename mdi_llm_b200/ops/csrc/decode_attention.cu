// Split-KV ("flash-decoding") attention for single-token decode over the slot KV pool, sm_100a.
//
// Reference behaviour being replaced: F.scaled_dot_product_attention over ALL S
// cache slots through a [1,1,1,S] boolean mask row, with K/V already expanded to n_head heads
// (model.py:704-751).  Here the cache holds the G group heads only, a CTA serves all q_per_kv
// query heads of one group from a single pass over that group's K/V (GQA packing: 4x fewer
// bytes for Llama-3), only the live prefix [0, pos] is scanned, and the sequence is split over
// `n_split` CTAs per group so the grid fills the GPU for any context length.
//
//   partial kernel : grid (G, n_split), 4 warps; each warp walks 32-position tiles with an
//                    online softmax per query head; CTA merges its warps and writes
//                    (m, l, acc[hs]) per (head, split).
//   combine kernel : grid (H), merges the splits -> bf16 y[H*hs] (input of the output proj).
#include "common.cuh"

namespace mdi {

constexpr int ATT_WARPS = 4;
constexpr int ATT_THREADS = ATT_WARPS * 32;
constexpr int ATT_TILE = 32;

struct AttnArgs {
  const bf16* q;   // [H * hs] (already roped)
  const bf16* kv;  // this layer's pool [n_slots, 2, G, S, hs]
  float* part;     // [H, n_split, hs + 2]
  const int* ctx;
  int n_head, n_groups, max_seq, n_split;
  float scale_log2;  // (1/sqrt(hs)) * log2(e)
};

template <int HS, int QPK>
__global__ void __launch_bounds__(ATT_THREADS) attn_decode_partial_kernel(const AttnArgs a) {
  constexpr int DPL = HS / 32;  // output dims per lane in the PV phase
  constexpr int QDIM = HS / 4;  // dims per lane in the QK phase (4 lanes per position)
  __shared__ __align__(16) float q_s[QPK][HS];
  __shared__ float s_s[ATT_WARPS][QPK][ATT_TILE];
  __shared__ float mrg_m[ATT_WARPS][QPK], mrg_l[ATT_WARPS][QPK];
  __shared__ float mrg_acc[ATT_WARPS][QPK][HS];

  const int g = blockIdx.x, split = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  pdl_wait_prior();
  const int slot = a.ctx[MDI_CTX_SLOT], L = a.ctx[MDI_CTX_POS] + 1;
  pdl_launch_dependents();

  // this split's position range, tile aligned
  const int tiles = (L + ATT_TILE - 1) / ATT_TILE;
  const int tiles_per_split = (tiles + a.n_split - 1) / a.n_split;
  const int t_lo = split * tiles_per_split, t_hi = min(tiles, t_lo + tiles_per_split);

  for (int i = threadIdx.x; i < QPK * HS; i += ATT_THREADS) {
    int h = i / HS, d = i % HS;
    q_s[h][d] = __bfloat162float(a.q[(size_t)(g * QPK + h) * HS + d]) * a.scale_log2;
  }
  __syncthreads();

  const bf16* kbase = a.kv + (((size_t)slot * 2 + 0) * a.n_groups + g) * (size_t)a.max_seq * HS;
  const bf16* vbase = a.kv + (((size_t)slot * 2 + 1) * a.n_groups + g) * (size_t)a.max_seq * HS;

  float m[QPK], l[QPK], acc[QPK][DPL];
#pragma unroll
  for (int h = 0; h < QPK; ++h) {
    m[h] = -INFINITY; l[h] = 0.f;
#pragma unroll
    for (int d = 0; d < DPL; ++d) acc[h][d] = 0.f;
  }

  for (int t = t_lo + warp; t < t_hi; t += ATT_WARPS) {
    const int p0 = t * ATT_TILE;
    // ---- QK^T: 4 lanes per position, 8 positions per pass, 4 passes -------------------------
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
      const int pj = pass * 8 + (lane >> 2), pos = p0 + pj, qd = (lane & 3) * QDIM;
      float sc[QPK];
#pragma unroll
      for (int h = 0; h < QPK; ++h) sc[h] = 0.f;
      if (pos < L) {
        const uint4* kr = reinterpret_cast<const uint4*>(kbase + (size_t)pos * HS + qd);
#pragma unroll
        for (int v = 0; v < QDIM / 8; ++v) {
          uint4 kk = __ldg(kr + v);
          float kf[8] = {bf16lo(kk.x), bf16hi(kk.x), bf16lo(kk.y), bf16hi(kk.y),
                         bf16lo(kk.z), bf16hi(kk.z), bf16lo(kk.w), bf16hi(kk.w)};
#pragma unroll
          for (int h = 0; h < QPK; ++h) {
            const float* qq = &q_s[h][qd + v * 8];
#pragma unroll
            for (int e = 0; e < 8; ++e) sc[h] = fmaf(kf[e], qq[e], sc[h]);
          }
        }
      }
#pragma unroll
      for (int h = 0; h < QPK; ++h) {
        float s = sc[h];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if ((lane & 3) == 0) s_s[warp][h][pj] = (pos < L) ? s : -INFINITY;
      }
    }
    __syncwarp();
    // ---- online softmax update: lane == position inside the tile --------------------------
#pragma unroll
    for (int h = 0; h < QPK; ++h) {
      const float s = s_s[warp][h][lane];
      const float m_new = fmaxf(m[h], warp_max(s));
      const float p = (s == -INFINITY) ? 0.f : exp2f(s - m_new);
      const float corr = (m[h] == -INFINITY) ? 0.f : exp2f(m[h] - m_new);
      l[h] = l[h] * corr + warp_sum(p);
      m[h] = m_new;
#pragma unroll
      for (int d = 0; d < DPL; ++d) acc[h][d] *= corr;
      s_s[warp][h][lane] = p;
    }
    __syncwarp();
    // ---- PV: lane owns DPL consecutive output dims --------------------------------------------
    const int n_pos = min(ATT_TILE, L - p0);
    for (int pj = 0; pj < n_pos; ++pj) {
      const bf16* vr = vbase + (size_t)(p0 + pj) * HS + lane * DPL;
      float vf[DPL];
      if (DPL == 4) {
        uint2 vv = __ldg(reinterpret_cast<const uint2*>(vr));
        vf[0] = bf16lo(vv.x); vf[1] = bf16hi(vv.x); vf[2 % DPL] = bf16lo(vv.y); vf[3 % DPL] = bf16hi(vv.y);
      } else {
#pragma unroll
        for (int d = 0; d < DPL; d += 2) {
          uint32_t vv = __ldg(reinterpret_cast<const uint32_t*>(vr + d));
          vf[d] = bf16lo(vv); vf[d + 1] = bf16hi(vv);
        }
      }
#pragma unroll
      for (int h = 0; h < QPK; ++h) {
        const float p = s_s[warp][h][pj];
#pragma unroll
        for (int d = 0; d < DPL; ++d) acc[h][d] = fmaf(p, vf[d], acc[h][d]);
      }
    }
    __syncwarp();
  }

  // ---- merge the CTA's warps ------------------------------------------------------------------
#pragma unroll
  for (int h = 0; h < QPK; ++h) {
    if (lane == 0) { mrg_m[warp][h] = m[h]; mrg_l[warp][h] = l[h]; }
#pragma unroll
    for (int d = 0; d < DPL; ++d) mrg_acc[warp][h][lane * DPL + d] = acc[h][d];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < QPK * HS; i += ATT_THREADS) {
    const int h = i / HS, d = i % HS;
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < ATT_WARPS; ++w) mm = fmaxf(mm, mrg_m[w][h]);
    float o = 0.f, ll = 0.f;
#pragma unroll
    for (int w = 0; w < ATT_WARPS; ++w) {
      const float sc = (mrg_m[w][h] == -INFINITY) ? 0.f : exp2f(mrg_m[w][h] - mm);
      o = fmaf(sc, mrg_acc[w][h][d], o);
      ll = fmaf(sc, mrg_l[w][h], ll);
    }
    float* dst = a.part + ((size_t)(g * QPK + h) * a.n_split + split) * (HS + 2);
    dst[2 + d] = o;
    if (d == 0) { dst[0] = mm; dst[1] = ll; }
  }
}

template <int HS>
__global__ void __launch_bounds__(HS) attn_decode_combine_kernel(const float* __restrict__ part, bf16* __restrict__ y,
                                                                 int n_split) {
  pdl_wait_prior();
  pdl_launch_dependents();
  const int h = blockIdx.x, d = threadIdx.x;
  const float* base = part + (size_t)h * n_split * (HS + 2);
  float mm = -INFINITY;
  for (int s = 0; s < n_split; ++s) mm = fmaxf(mm, base[(size_t)s * (HS + 2)]);
  float o = 0.f, ll = 0.f;
  for (int s = 0; s < n_split; ++s) {
    const float* p = base + (size_t)s * (HS + 2);
    const float sc = (p[0] == -INFINITY) ? 0.f : exp2f(p[0] - mm);
    o = fmaf(sc, p[2 + d], o);
    ll = fmaf(sc, p[1], ll);
  }
  y[(size_t)h * HS + d] = __float2bfloat16_rn(ll > 0.f ? o / ll : 0.f);
}

template <int HS, int QPK>
static int launch_attn(const AttnArgs& a, bf16* y, int use_pdl, cudaStream_t stream) {
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(a.n_groups, a.n_split);
  cfg.blockDim = dim3(ATT_THREADS);
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  cudaError_t e = cudaLaunchKernelEx(&cfg, attn_decode_partial_kernel<HS, QPK>, a);
  if (e != cudaSuccess) return (int)e;
  cfg.gridDim = dim3(a.n_head);
  cfg.blockDim = dim3(HS);
  return (int)cudaLaunchKernelEx(&cfg, attn_decode_combine_kernel<HS>, (const float*)a.part, y, a.n_split);
}

}  // namespace mdi

using namespace mdi;

extern "C" int mdi_attn_decode(const void* q, const void* kv, void* y, float* part, const int* ctx, int n_head,
                               int n_groups, int head_size, int max_seq, int n_split, int use_pdl,
                               cudaStream_t stream) {
  AttnArgs a;
  a.q = (const bf16*)q; a.kv = (const bf16*)kv; a.part = part; a.ctx = ctx;
  a.n_head = n_head; a.n_groups = n_groups; a.max_seq = max_seq; a.n_split = n_split;
  a.scale_log2 = 1.4426950408889634f / sqrtf((float)head_size);
  const int qpk = n_head / n_groups;
#define MDI_ATT_CASE(HS_, QPK_) \
  if (head_size == HS_ && qpk == QPK_) return launch_attn<HS_, QPK_>(a, (bf16*)y, use_pdl, stream);
  MDI_ATT_CASE(128, 1) MDI_ATT_CASE(128, 2) MDI_ATT_CASE(128, 4) MDI_ATT_CASE(128, 8)
  MDI_ATT_CASE(64, 1) MDI_ATT_CASE(64, 2) MDI_ATT_CASE(64, 4) MDI_ATT_CASE(64, 8)
#undef MDI_ATT_CASE
  return -3;  // unsupported (head_size, q_per_kv): caller falls back to the eager path
}

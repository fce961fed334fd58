// Split-KV ("flash-decoding") attention for single-token decode over the slot KV pool, sm_100a.
//
// Reference behaviour being replaced: F.scaled_dot_product_attention over ALL S
// cache slots through a [1,1,1,S] boolean mask row, with K/V already expanded to n_head heads
// (model.py:704-751).  Here the cache holds the G group heads only, a CTA serves all q_per_kv
// query heads of one group from a single pass over that group's K/V (GQA packing: 4x fewer
// bytes for Llama-3) and only the live prefix [0, pos] is scanned.
//
// One kernel, grid (G, n_split), 8 warps per CTA, a warp owns 16-position tiles:
//   * the context is cut into spans of >= 128 positions, so short contexts use few CTAs and
//     the others exit at once (the grid is fixed at graph-capture time, the length is not);
//   * ALL K and V loads of a tile are issued together before anything is consumed — decode attention
//     is a latency chain (ctx -> q -> K -> softmax -> V -> merge), not a bandwidth problem; small
//     tiles over many warps keep the serial instruction count per warp low;
//   * the last CTA of a group to finish merges the spans (atomic ticket) and writes bf16
//     y[H*hs] directly: no separate combine launch;
//   * K / V of OLD positions do not depend on this step's QKV projection: every warp requests its first tile
//     BEFORE the programmatic-dependency wait (the kernel is resident while the QKV kernel still streams its
//     weights), so after the wait only q — and the tile that holds the newest position — remain to be fetched.
#include "common.cuh"

namespace mdi {

constexpr int ATT_WARPS = 8;
constexpr int ATT_THREADS = ATT_WARPS * 32;
constexpr int ATT_TILE = 16;
constexpr int ATT_NPASS = ATT_TILE / 8;  // QK passes per tile (4 lanes per position, 8 positions per pass)

struct AttnArgs {
  const bf16* q;   // [H * hs] (already roped)
  const bf16* kv;  // this layer's pool [n_slots, 2, G, S, hs]
  bf16* y;         // [H * hs]
  float* part;     // [H, n_split, hs + 2]
  unsigned int* tickets;  // [G] zero-initialised, self-resetting
  const int* ctx;
  int n_head, n_groups, max_seq, n_split;
  float scale_log2;  // (1/sqrt(hs)) * log2(e)
  unsigned long long* trace;
  DepWait dep_wait;      // flag dependency on the QKV kernel (else griddepcontrol.wait)
  DepSignal dep_signal;  // flag for the out-projection
};

// CLUSTER: the spans of a KV group are launched as thread-block clusters of ATT_CL CTAs along y.  While the live
// context fits into one cluster (n_active <= ATT_CL spans, i.e. up to 1024 positions at 128 per span) the spans are
// merged through DISTRIBUTED SHARED MEMORY: every CTA writes its (max, sum, partial output) into rank 0's shared
// memory (st.shared::cluster), one cluster barrier, rank 0 folds them and writes y — no partial buffer in global
// memory, no fence + ticket, no second round of L2 reads (measured: the global merge was 3.8 us of the kernel's 11).
// Longer contexts fall back to the ticket merge below.
constexpr int ATT_CL = 8;

__device__ __forceinline__ uint32_t att_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void st_cluster_f32(const float* local_smem, uint32_t cta, float v) {
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(smem_u32(local_smem)), "r"(cta));
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
}

template <int HS, int QPK, bool CLUSTER>
__global__ void __launch_bounds__(ATT_THREADS) attn_decode_kernel(const AttnArgs a) {
  constexpr int DPL = HS / 32;   // output dims per lane in the PV phase
  constexpr int QDIM = HS / 4;   // dims per lane in the QK phase (4 lanes per position)
  constexpr int KV4 = QDIM / 8;  // uint4 loads per lane per position
  constexpr int VB = (QPK >= 8 && HS >= 128) ? 8 : ATT_TILE;  // V rows per load batch (register budget)
  constexpr int KPASS = (QPK >= 8 && HS >= 128) ? 1 : ATT_NPASS;  // K passes whose loads are batched
  constexpr bool V_EARLY = (VB == ATT_TILE);  // the whole V tile is requested together with the K tile
  __shared__ __align__(16) float q_s[QPK][HS];
  __shared__ float s_s[ATT_WARPS][QPK][ATT_TILE];
  __shared__ float mrg_m[ATT_WARPS][QPK], mrg_l[ATT_WARPS][QPK];
  __shared__ float mrg_acc[ATT_WARPS][QPK][HS];
  __shared__ int sh_last;
  __shared__ float cl_part[CLUSTER ? ATT_CL : 1][CLUSTER ? QPK : 1][CLUSTER ? HS + 2 : 1];  // rank 0: the cluster's partials

  const int g = blockIdx.x, split = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  trace_mark(a.trace, 0, true);
  // ctx was written at the start of the step (advance_step / the host's descriptor copy), at least two launches ago:
  // readable before the dependency wait
  const int slot = a.ctx[MDI_CTX_SLOT], L = a.ctx[MDI_CTX_POS] + 1;
  // spans of whole tiles, at least ATT_WARPS tiles (128 positions) each
  const int tiles = (L + ATT_TILE - 1) / ATT_TILE;
  int tiles_per_split = (tiles + a.n_split - 1) / a.n_split;
  if (tiles_per_split < ATT_WARPS) tiles_per_split = ATT_WARPS;
  const int n_active = (tiles + tiles_per_split - 1) / tiles_per_split;
  const bool active = split < n_active;
  const int t_lo = split * tiles_per_split, t_hi = min(tiles, t_lo + tiles_per_split);
  const bf16* kbase = a.kv + (((size_t)slot * 2 + 0) * a.n_groups + g) * (size_t)a.max_seq * HS;
  const bf16* vbase = a.kv + (((size_t)slot * 2 + 1) * a.n_groups + g) * (size_t)a.max_seq * HS;

  // K / V registers of one tile.  PRE: the whole tile's K (all passes) and V are requested together, so the warp's
  // first tile can be requested before the wait when it only holds positions older than this step's
  constexpr bool PRE = V_EARLY && (KPASS == ATT_NPASS);
  uint4 kk[KPASS][KV4];
  uint32_t vv[VB][DPL / 2];
  auto load_k = [&](int p0, int pg) {
#pragma unroll
    for (int pp = 0; pp < KPASS; ++pp) {
      const int pos = min(p0 + (pg * KPASS + pp) * 8 + (lane >> 2), L - 1);  // clamp: masked below
      const uint4* kr = reinterpret_cast<const uint4*>(kbase + (size_t)pos * HS + (lane & 3) * QDIM);
#pragma unroll
      for (int v = 0; v < KV4; ++v) kk[pp][v] = __ldg(kr + v);
    }
  };
  auto load_v = [&](int p0, int half) {
#pragma unroll
    for (int r = 0; r < VB; ++r) {
      const int pos = min(p0 + half * VB + r, L - 1);  // masked rows have p == 0
      const uint32_t* vr = reinterpret_cast<const uint32_t*>(vbase + (size_t)pos * HS + lane * DPL);
#pragma unroll
      for (int w = 0; w < DPL / 2; ++w) vv[r][w] = __ldg(vr + w);
    }
  };
  const int t_first = t_lo + warp;
  bool pre = false;
  if (PRE && active && t_first < t_hi && (t_first + 1) * ATT_TILE <= L - 1) {  // no row of this tile is written by this step
    load_k(t_first * ATT_TILE, 0);
    load_v(t_first * ATT_TILE, 0);
    pre = true;
  }
  if (a.dep_wait.flag) dep_wait(a.dep_wait, a.ctx); else pdl_wait_prior();
  trace_mark(a.trace, 1, true);
  pdl_launch_dependents();
  // merge through distributed shared memory: every live span sits in the group's first cluster
  const bool cl_merge = CLUSTER && n_active > 1 && n_active <= ATT_CL;
  if (!active) {  // idle span: only its ticket (and, inside the merging cluster, its barrier arrival)
    if (cl_merge && split < ATT_CL) asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    dep_signal(a.dep_signal, a.ctx);
    return;
  }
  trace_mark(a.trace, 2, false);

  for (int i = threadIdx.x; i < QPK * HS; i += ATT_THREADS) {
    const int h = i / HS, d = i % HS;
    q_s[h][d] = __bfloat162float(a.q[(size_t)(g * QPK + h) * HS + d]) * a.scale_log2;
  }
  __syncthreads();

  float m[QPK], l[QPK], acc[QPK][DPL];
#pragma unroll
  for (int h = 0; h < QPK; ++h) {
    m[h] = -INFINITY; l[h] = 0.f;
#pragma unroll
    for (int d = 0; d < DPL; ++d) acc[h][d] = 0.f;
  }

  for (int t = t_first; t < t_hi; t += ATT_WARPS) {
    const int p0 = t * ATT_TILE;
    const bool have = PRE && pre && t == t_first;  // requested before the wait
    // ---- issue the K loads of KPASS passes (8 positions each) and the V rows up-front, then consume --
#pragma unroll
    for (int pg = 0; pg < ATT_NPASS / KPASS; ++pg) {
      if (!have) {
        load_k(p0, pg);
        if (V_EARLY && pg == 0) load_v(p0, 0);
      }
#pragma unroll
      for (int pp = 0; pp < KPASS; ++pp) {
        const int pj = (pg * KPASS + pp) * 8 + (lane >> 2), pos = p0 + pj, qd = (lane & 3) * QDIM;
        float sc[QPK];
#pragma unroll
        for (int h = 0; h < QPK; ++h) sc[h] = 0.f;
#pragma unroll
        for (int v = 0; v < KV4; ++v) {
          const uint4 k4 = kk[pp][v];
          const float kf[8] = {bf16lo(k4.x), bf16hi(k4.x), bf16lo(k4.y), bf16hi(k4.y),
                               bf16lo(k4.z), bf16hi(k4.z), bf16lo(k4.w), bf16hi(k4.w)};
#pragma unroll
          for (int h = 0; h < QPK; ++h) {
            const float4 qa = *reinterpret_cast<const float4*>(&q_s[h][qd + v * 8]);
            const float4 qb = *reinterpret_cast<const float4*>(&q_s[h][qd + v * 8 + 4]);
            sc[h] = fmaf(kf[0], qa.x, sc[h]); sc[h] = fmaf(kf[1], qa.y, sc[h]);
            sc[h] = fmaf(kf[2], qa.z, sc[h]); sc[h] = fmaf(kf[3], qa.w, sc[h]);
            sc[h] = fmaf(kf[4], qb.x, sc[h]); sc[h] = fmaf(kf[5], qb.y, sc[h]);
            sc[h] = fmaf(kf[6], qb.z, sc[h]); sc[h] = fmaf(kf[7], qb.w, sc[h]);
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
    }
    __syncwarp();
    // ---- online softmax update: lane == position inside the tile --------------------------
#pragma unroll
    for (int h = 0; h < QPK; ++h) {
      const float s = lane < ATT_TILE ? s_s[warp][h][lane] : -INFINITY;
      const float m_new = fmaxf(m[h], warp_max(s));
      const float p = (s == -INFINITY) ? 0.f : exp2f(s - m_new);
      const float corr = (m[h] == -INFINITY) ? 0.f : exp2f(m[h] - m_new);
      l[h] = l[h] * corr + warp_sum(p);
      m[h] = m_new;
#pragma unroll
      for (int d = 0; d < DPL; ++d) acc[h][d] *= corr;
      if (lane < ATT_TILE) s_s[warp][h][lane] = p;
    }
    __syncwarp();
    // ---- PV: lane owns DPL consecutive output dims ---------------------------------------------------
#pragma unroll
    for (int half = 0; half < ATT_TILE / VB; ++half) {
      if (!V_EARLY) load_v(p0, half);
#pragma unroll
      for (int r = 0; r < VB; ++r) {
#pragma unroll
        for (int h = 0; h < QPK; ++h) {
          const float p = s_s[warp][h][half * VB + r];
#pragma unroll
          for (int w = 0; w < DPL / 2; ++w) {
            acc[h][2 * w] = fmaf(p, bf16lo(vv[r][w]), acc[h][2 * w]);
            acc[h][2 * w + 1] = fmaf(p, bf16hi(vv[r][w]), acc[h][2 * w + 1]);
          }
        }
      }
    }
    __syncwarp();
  }

  __syncthreads();
  trace_phase(a.trace, 6);  // all tiles done
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
    if (n_active == 1) {  // single span: final answer, no round trip through the partial buffer
      a.y[(size_t)(g * QPK + h) * HS + d] = __float2bfloat16_rn(ll > 0.f ? o / ll : 0.f);
    } else if (cl_merge) {  // into rank 0's shared memory (my slot = my span index)
      st_cluster_f32(&cl_part[split][h][2 + d], 0, o);
      if (d == 0) { st_cluster_f32(&cl_part[split][h][0], 0, mm); st_cluster_f32(&cl_part[split][h][1], 0, ll); }
    } else {
      float* dst = a.part + ((size_t)(g * QPK + h) * a.n_split + split) * (HS + 2);
      dst[2 + d] = o;
      if (d == 0) { dst[0] = mm; dst[1] = ll; }
    }
  }
  if (n_active == 1) { dep_signal(a.dep_signal, a.ctx); trace_mark(a.trace, 3, true); trace_mark(a.trace, 4, false); return; }

  if (CLUSTER && cl_merge) {
    // ---- rank 0 folds the cluster's partials out of its own shared memory ------------------------------
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    if (split != 0) { dep_signal(a.dep_signal, a.ctx); trace_mark(a.trace, 3, true); trace_mark(a.trace, 4, false); return; }
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    for (int i = threadIdx.x; i < QPK * HS; i += ATT_THREADS) {
      const int h = i / HS, d = i % HS;
      float mm = -INFINITY;
      for (int sp = 0; sp < n_active; ++sp) mm = fmaxf(mm, cl_part[sp][h][0]);
      float o = 0.f, ll = 0.f;
      for (int sp = 0; sp < n_active; ++sp) {
        const float pm = cl_part[sp][h][0];
        const float sc = (pm == -INFINITY) ? 0.f : exp2f(pm - mm);
        o = fmaf(sc, cl_part[sp][h][2 + d], o);
        ll = fmaf(sc, cl_part[sp][h][1], ll);
      }
      a.y[(size_t)(g * QPK + h) * HS + d] = __float2bfloat16_rn(ll > 0.f ? o / ll : 0.f);
    }
    dep_signal(a.dep_signal, a.ctx);
    trace_mark(a.trace, 3, true);
    trace_mark(a.trace, 4, false);
    return;
  }

  // ---- the group's last CTA merges the spans ------------------------------------------------------
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(a.tickets + g, 1u);
    sh_last = (prev == (unsigned)n_active - 1);
    if (sh_last) a.tickets[g] = 0;
  }
  __syncthreads();
  if (!sh_last) { dep_signal(a.dep_signal, a.ctx); trace_mark(a.trace, 3, true); trace_mark(a.trace, 4, false); return; }
  __threadfence();
  for (int i = threadIdx.x; i < QPK * HS; i += ATT_THREADS) {
    const int h = i / HS, d = i % HS;
    const float* base = a.part + (size_t)(g * QPK + h) * a.n_split * (HS + 2);
    float mm = -INFINITY;
    for (int s = 0; s < n_active; ++s) mm = fmaxf(mm, __ldcg(base + (size_t)s * (HS + 2)));
    float o = 0.f, ll = 0.f;
#pragma unroll 4
    for (int s = 0; s < n_active; ++s) {
      const float* p = base + (size_t)s * (HS + 2);
      const float pm = __ldcg(p), pl = __ldcg(p + 1), po = __ldcg(p + 2 + d);
      const float sc = (pm == -INFINITY) ? 0.f : exp2f(pm - mm);
      o = fmaf(sc, po, o);
      ll = fmaf(sc, pl, ll);
    }
    a.y[(size_t)(g * QPK + h) * HS + d] = __float2bfloat16_rn(ll > 0.f ? o / ll : 0.f);
  }
  dep_signal(a.dep_signal, a.ctx);
  trace_mark(a.trace, 3, true);
  trace_mark(a.trace, 4, false);
}

static int g_attn_cluster = 1;  // mdi_set_attn_cluster: 0 = always merge through global memory

template <int HS, int QPK>
static int launch_attn(const AttnArgs& a, int use_pdl, cudaStream_t stream) {
  // the cluster's partials + the kernel's other static arrays must fit the 48 KB static shared-memory window
  constexpr bool CL_FITS = (size_t)ATT_CL * QPK * (HS + 2) * 4 + (size_t)ATT_WARPS * QPK * HS * 4 + (size_t)QPK * HS * 4 + 4096 <= 48 * 1024;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(a.n_groups, a.n_split);
  cfg.blockDim = dim3(ATT_THREADS);
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  if (CL_FITS && g_attn_cluster && a.n_split % ATT_CL == 0) {
    if (!use_pdl) { attr[0].id = cudaLaunchAttributeClusterDimension; attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = ATT_CL; attr[0].val.clusterDim.z = 1; }
    else { attr[1].id = cudaLaunchAttributeClusterDimension; attr[1].val.clusterDim.x = 1; attr[1].val.clusterDim.y = ATT_CL; attr[1].val.clusterDim.z = 1; }
    cfg.numAttrs += 1;
    return (int)cudaLaunchKernelEx(&cfg, attn_decode_kernel<HS, QPK, CL_FITS>, a);
  }
  return (int)cudaLaunchKernelEx(&cfg, attn_decode_kernel<HS, QPK, false>, a);
}

}  // namespace mdi

using namespace mdi;

extern "C" void mdi_set_attn_cluster(int on) { g_attn_cluster = on; }

// part: fp32 [H, n_split, hs + 2]; tickets: uint32 [G], zeroed once at allocation.
extern "C" int mdi_attn_decode(const void* q, const void* kv, void* y, float* part, unsigned int* tickets,
                               const int* ctx, int n_head, int n_groups, int head_size, int max_seq, int n_split,
                               int use_pdl, unsigned long long* trace, const int* dep_wait_flag, int* dep_signal_flag,
                               unsigned int* dep_ctr, int* status, long long wait_max_cycles, cudaStream_t stream) {
  AttnArgs a;
  a.trace = trace;
  a.dep_wait = DepWait{dep_wait_flag, status, wait_max_cycles}; a.dep_signal = DepSignal{dep_signal_flag, dep_ctr};
  a.q = (const bf16*)q; a.kv = (const bf16*)kv; a.y = (bf16*)y; a.part = part; a.tickets = tickets; a.ctx = ctx;
  a.n_head = n_head; a.n_groups = n_groups; a.max_seq = max_seq; a.n_split = n_split;
  a.scale_log2 = 1.4426950408889634f / sqrtf((float)head_size);
  const int qpk = n_head / n_groups;
#define MDI_ATT_CASE(HS_, QPK_) \
  if (head_size == HS_ && qpk == QPK_) return launch_attn<HS_, QPK_>(a, use_pdl, stream);
  MDI_ATT_CASE(128, 1) MDI_ATT_CASE(128, 2) MDI_ATT_CASE(128, 4) MDI_ATT_CASE(128, 8)
  MDI_ATT_CASE(64, 1) MDI_ATT_CASE(64, 2) MDI_ATT_CASE(64, 4) MDI_ATT_CASE(64, 8)
  MDI_ATT_CASE(256, 1) MDI_ATT_CASE(256, 2)  // Gemma-7b / Pythia-1b class heads (static shared memory <= 48 KB)
#undef MDI_ATT_CASE
  return -3;  // unsupported (head_size, q_per_kv): caller falls back to the eager path
}

// Weight-streaming linear layers for single-token decode (M = 1), sm_100a.
//
// Every decode linear of a litGPT block is HBM-bandwidth bound (one
// pass over the weight matrix per token), so the kernels here are organised around keeping
// enough 16-byte weight loads in flight (8 per lane, 8 warps, several CTAs per SM) and fusing
// everything else into that single pass:
//
//   prologue : hop wait (acquire the previous stage's flag), RMSNorm of the input row
//              (replaces the reference's 5+ elementwise launches, model.py:966-977)
//   epilogue : bias, SiLU/GELU gating (model.py:805-820), residual add (model.py:625-628),
//              RoPE + KV-cache append for the QKV projection (model.py:693-729,931-932),
//              P2P store of the hidden state into the next stage's buffer + release flag
//              (replaces pickle+TCP, connections.py:325-353).
//
// A warp computes two output rows at a time ("row pair"): for gated MLPs the pair is
// (fc_1[n], fc_2[n]); for QKV it is the two rows that RoPE rotates together; otherwise two
// adjacent rows.  The activation vector lives in shared memory as bf16.
#include <cuda_fp16.h>
#include <cuda_fp8.h>

#include "common.cuh"

namespace mdi {

constexpr int LIN_THREADS = 256;
constexpr int LIN_WARPS = LIN_THREADS / 32;
constexpr int LIN_UNROLL = 4;

__device__ __forceinline__ float dot8(const uint4& w, const uint4& x) {
  float s = bf16lo(w.x) * bf16lo(x.x);
  s = fmaf(bf16hi(w.x), bf16hi(x.x), s);
  s = fmaf(bf16lo(w.y), bf16lo(x.y), s);
  s = fmaf(bf16hi(w.y), bf16hi(x.y), s);
  s = fmaf(bf16lo(w.z), bf16lo(x.z), s);
  s = fmaf(bf16hi(w.z), bf16hi(x.z), s);
  s = fmaf(bf16lo(w.w), bf16lo(x.w), s);
  s = fmaf(bf16hi(w.w), bf16hi(x.w), s);
  return s;
}

// Dot products of two weight rows with the shared-memory activation vector (nvec = K / 8).
__device__ __forceinline__ void warp_dot2(const bf16* __restrict__ wa, const bf16* __restrict__ wb,
                                          const uint4* __restrict__ xs, int nvec, int lane,
                                          float& out_a, float& out_b) {
  const uint4* pa = reinterpret_cast<const uint4*>(wa);
  const uint4* pb = reinterpret_cast<const uint4*>(wb);
  float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
  int v = lane;
  for (; v + 32 * (LIN_UNROLL - 1) < nvec; v += 32 * LIN_UNROLL) {
    uint4 ra[LIN_UNROLL], rb[LIN_UNROLL];
#pragma unroll
    for (int u = 0; u < LIN_UNROLL; ++u) {
      ra[u] = ldg_stream(pa + v + 32 * u);
      rb[u] = ldg_stream(pb + v + 32 * u);
    }
#pragma unroll
    for (int u = 0; u < LIN_UNROLL; ++u) {
      uint4 x = xs[v + 32 * u];
      if (u & 1) {
        a1 += dot8(ra[u], x);
        b1 += dot8(rb[u], x);
      } else {
        a0 += dot8(ra[u], x);
        b0 += dot8(rb[u], x);
      }
    }
  }
  for (; v < nvec; v += 32) {
    uint4 x = xs[v];
    a0 += dot8(ldg_stream(pa + v), x);
    b0 += dot8(ldg_stream(pb + v), x);
  }
  out_a = warp_sum(a0 + a1);
  out_b = warp_sum(b0 + b1);
}

// Stage the activation row into shared memory, optionally RMS-normalised:
//   xn = bf16(x * rsqrt(mean(x^2) + eps)) * (w | 1 + w)      (same rounding points as the eager model)
// Fully vectorised: every thread owns up to STAGE_VPT 16-byte vectors of x (and of the norm weight),
// all loads are issued before the first use, the row stays in registers across the reduction.
// (The first version re-read x from shared memory and fetched the norm weight with 2-byte global
//  loads inside the loop: 16 dependent L2 round trips = 11 us on the QKV kernel — see
//  profiles/trace_step_*.json.)
constexpr int STAGE_VPT = 4;  // covers K <= 4 * 256 * 8 = 8192 (n_embd of 70B models); longer rows use the loop

__device__ __forceinline__ uint32_t norm_pair(uint32_t xp, uint32_t wp, float rstd, int unit_offset) {
  float x0 = round_bf16(bf16lo(xp) * rstd), x1 = round_bf16(bf16hi(xp) * rstd);
  float w0 = bf16lo(wp), w1 = bf16hi(wp);
  if (unit_offset) { w0 = round_bf16(1.f + w0); w1 = round_bf16(1.f + w1); }
  const __nv_bfloat162 o = __floats2bfloat162_rn(x0 * w0, x1 * w1);
  return *reinterpret_cast<const uint32_t*>(&o);
}

// LayerNorm variant (GPT-2 / Pythia / Falcon / Phi families): y = bf16((x - mean) * rstd * w + b), statistics in
// fp32 — what torch's layer_norm computes on a bf16 row (one rounding at the end).
__device__ __forceinline__ uint32_t ln_pair(uint32_t xp, uint32_t wp, uint32_t bp, float mean, float rstd, bool has_b) {
  const float y0 = (bf16lo(xp) - mean) * rstd * bf16lo(wp) + (has_b ? bf16lo(bp) : 0.f);
  const float y1 = (bf16hi(xp) - mean) * rstd * bf16hi(wp) + (has_b ? bf16hi(bp) : 0.f);
  const __nv_bfloat162 o = __floats2bfloat162_rn(y0, y1);
  return *reinterpret_cast<const uint32_t*>(&o);
}

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();  // red may still be read from a previous reduction
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < LIN_THREADS / 32; ++w) tot += red[w];
  return tot;
}

__device__ __forceinline__ void stage_input_layernorm(const bf16* __restrict__ x, const bf16* __restrict__ norm_w,
                                                      const bf16* __restrict__ norm_b, float eps, int K, bf16* xs, float* red) {
  const int tid = threadIdx.x;
  const int nvec = K / 8;
  const uint4* src = reinterpret_cast<const uint4*>(x);
  const uint4* wsrc = reinterpret_cast<const uint4*>(norm_w);
  const uint4* bsrc = reinterpret_cast<const uint4*>(norm_b);
  uint4* dst = reinterpret_cast<uint4*>(xs);
  float s = 0.f;
  for (int v = tid; v < nvec; v += LIN_THREADS) {  // pass 1: row to shared memory + sum
    const uint4 r = __ldcg(src + v);
    dst[v] = r;
    const uint32_t p[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) s += bf16lo(p[q]) + bf16hi(p[q]);
  }
  const float mean = block_sum(s, red) / (float)K;
  float ss = 0.f;
  for (int v = tid; v < nvec; v += LIN_THREADS) {  // pass 2: centred second moment (own vectors: no sync needed)
    const uint4 r = dst[v];
    const uint32_t p[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) { const float a = bf16lo(p[q]) - mean, b = bf16hi(p[q]) - mean; ss = fmaf(a, a, ss); ss = fmaf(b, b, ss); }
  }
  const float rstd = rsqrtf(block_sum(ss, red) / (float)K + eps);
  for (int v = tid; v < nvec; v += LIN_THREADS) {
    const uint4 r = dst[v], wv = __ldg(wsrc + v);
    const uint4 bv = norm_b ? __ldg(bsrc + v) : make_uint4(0, 0, 0, 0);
    uint4 o;
    o.x = ln_pair(r.x, wv.x, bv.x, mean, rstd, norm_b != nullptr); o.y = ln_pair(r.y, wv.y, bv.y, mean, rstd, norm_b != nullptr);
    o.z = ln_pair(r.z, wv.z, bv.z, mean, rstd, norm_b != nullptr); o.w = ln_pair(r.w, wv.w, bv.w, mean, rstd, norm_b != nullptr);
    dst[v] = o;
  }
  __syncthreads();
}

// `w_pre`: the norm weights of this thread's vectors, requested BEFORE the dependency wait (they are static; after
// ~200 MB of streamed matrices they are no longer in L2, so loading them next to x put a DRAM round trip on the
// post-wait critical path of every normalising kernel).
__device__ __forceinline__ void preload_norm_w(const bf16* __restrict__ norm_w, int K, uint4 (&w_pre)[4]) {
  const uint4* wsrc = reinterpret_cast<const uint4*>(norm_w);
  const int nvec = K / 8;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int v = threadIdx.x + j * 256;
    w_pre[j] = v < nvec ? __ldg(wsrc + v) : make_uint4(0, 0, 0, 0);
  }
}

__device__ __forceinline__ void stage_input(const bf16* __restrict__ x, const bf16* __restrict__ norm_w,
                                            float eps, int unit_offset, int K, bf16* xs, float* red,
                                            const bf16* __restrict__ norm_b = nullptr, int layer_norm = 0,
                                            const uint4* w_pre = nullptr) {
  if (layer_norm && norm_w != nullptr) { stage_input_layernorm(x, norm_w, norm_b, eps, K, xs, red); return; }
  const int tid = threadIdx.x;
  const int nvec = K / 8;
  const uint4* src = reinterpret_cast<const uint4*>(x);
  uint4* dst = reinterpret_cast<uint4*>(xs);
  if (norm_w == nullptr) {  // plain copy: all loads in flight before the first store
    for (int v0 = 0; v0 < nvec; v0 += 8 * LIN_THREADS) {
      uint4 r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int v = v0 + tid + j * LIN_THREADS;
        if (v < nvec) r[j] = __ldcg(src + v);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int v = v0 + tid + j * LIN_THREADS;
        if (v < nvec) dst[v] = r[j];
      }
    }
    __syncthreads();
    return;
  }
  const uint4* wsrc = reinterpret_cast<const uint4*>(norm_w);
  if (nvec <= STAGE_VPT * LIN_THREADS) {
    uint4 xr[STAGE_VPT], wr[STAGE_VPT];
#pragma unroll
    for (int j = 0; j < STAGE_VPT; ++j) {
      const int v = tid + j * LIN_THREADS;
      if (v < nvec) { xr[j] = __ldcg(src + v); wr[j] = w_pre ? w_pre[j] : __ldg(wsrc + v); }
      else { xr[j] = make_uint4(0, 0, 0, 0); wr[j] = xr[j]; }
    }
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < STAGE_VPT; ++j) {
      const uint32_t p[4] = {xr[j].x, xr[j].y, xr[j].z, xr[j].w};
#pragma unroll
      for (int q = 0; q < 4; ++q) { const float a = bf16lo(p[q]), b = bf16hi(p[q]); ss = fmaf(a, a, ss); ss = fmaf(b, b, ss); }
    }
    ss = warp_sum(ss);
    if ((tid & 31) == 0) red[tid >> 5] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < LIN_WARPS; ++w) tot += red[w];
    const float rstd = rsqrtf(tot / (float)K + eps);
#pragma unroll
    for (int j = 0; j < STAGE_VPT; ++j) {
      const int v = tid + j * LIN_THREADS;
      if (v < nvec) {
        uint4 o;
        o.x = norm_pair(xr[j].x, wr[j].x, rstd, unit_offset);
        o.y = norm_pair(xr[j].y, wr[j].y, rstd, unit_offset);
        o.z = norm_pair(xr[j].z, wr[j].z, rstd, unit_offset);
        o.w = norm_pair(xr[j].w, wr[j].w, rstd, unit_offset);
        dst[v] = o;
      }
    }
    __syncthreads();
    return;
  }
  // very long rows: two vectorised passes through shared memory
  float ss = 0.f;
  for (int v = tid; v < nvec; v += LIN_THREADS) {
    const uint4 r = __ldcg(src + v);
    dst[v] = r;
    const uint32_t p[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) { const float a = bf16lo(p[q]), b = bf16hi(p[q]); ss = fmaf(a, a, ss); ss = fmaf(b, b, ss); }
  }
  ss = warp_sum(ss);
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < LIN_WARPS; ++w) tot += red[w];
  const float rstd = rsqrtf(tot / (float)K + eps);
  for (int v = tid; v < nvec; v += LIN_THREADS) {
    const uint4 r = dst[v], wv = __ldg(wsrc + v);
    uint4 o;
    o.x = norm_pair(r.x, wv.x, rstd, unit_offset); o.y = norm_pair(r.y, wv.y, rstd, unit_offset);
    o.z = norm_pair(r.z, wv.z, rstd, unit_offset); o.w = norm_pair(r.w, wv.w, rstd, unit_offset);
    dst[v] = o;
  }
  __syncthreads();
}

// One argument block for every weight-streaming decode linear (plain / gated / QKV).
struct StreamArgs {
  const bf16* W;         // [N, K]   (QKV: [(H + 2G) * hs, K], litGPT group-interleaved rows)
  const bf16* W2;        // gated: second projection [N, K]; else null
  const bf16* bias;      // [N] or null
  const bf16* bias2;     // [N] or null
  const bf16* x;         // input activation row: x + slot * x_slot_stride
  const bf16* norm_w;    // fused RMSNorm / LayerNorm weight [K] or null
  const bf16* norm_b;    // LayerNorm bias [K] or null
  int layer_norm;        // 1: LayerNorm statistics (mean / variance), 0: RMS
  const bf16* residual;  // residual + slot * res_slot_stride, [N], or null
  void* y;               // output (bf16, or fp32 when out_fp32): y + slot * y_slot_stride
  const int* ctx;
  long long x_slot_stride, res_slot_stride, y_slot_stride;  // in elements; 0 = not slotted
  int N, K;
  float eps;
  int unit_offset;
  int act;
  int out_fp32;
  int n_items;
  HopWait wait;
  HopSignal signal;
  // QKV-only
  const float* cos;  // [S, n_elem]
  const float* sin;
  bf16* q_out;       // [H * hs]
  bf16* kv;          // this layer's pool: [n_slots, 2, G, S, hs]
  int n_head, n_groups, head_size, rope_n_elem, max_seq;
  // logits-only (lm_head): sampling statistics gathered while the logits are produced
  unsigned int* hist;        // [4096] histogram of the top 12 bits of the orderable logit key, or null
  unsigned long long* amax;  // packed (key << 32 | ~row) running arg-max, or null
  unsigned long long* trace; // tracer record of this launch (6 x u64) or null
  // fp8 (e4m3) block-scaled weights: W/W2 then point at 1-byte elements and these hold one fp32
  // scale per 128 consecutive K elements of every row ([N, K/128]); null = bf16 weights
  const float* wscale;
  const float* wscale2;
  DepWait dep_wait;      // intra-stage dependency on the previous kernel by flag (else griddepcontrol.wait)
  DepSignal dep_signal;  // ... and the flag this kernel publishes for the next one
  int l2_pf_chunks;  // bulk kernels: chunk pairs per warp prefetched into L2 beyond the smem ring (0 = off)
  int ctx_early;     // ctx was written >= 2 launches ago: slot/pos may be read before the PDL wait
  // hop by row copy: y is a LOCAL row buffer; the last CTA copies it to hop_row (+ slot * hop_slot_stride) in
  // the next stage's memory and releases the flag (hop_signal_copy); null = epilogues store to y directly
  bf16* hop_row;
  long long hop_slot_stride;
  const bf16* hop_pre;            // [x | h] messages: the residual row copied in front (hop_pre + slot * stride)
  long long hop_pre_slot_stride;
  int hop_pre_elems;
  // look-ahead for the NEXT kernel: while this kernel waits for its input (o_proj behind the latency-bound
  // attention: HBM idle) its warps ask the TMA engine to pull `pf_bytes` of each of these regions into L2
  const unsigned char* pf_a;
  const unsigned char* pf_b;
  unsigned long long pf_bytes;
};

// bulk L2 prefetch of [base, base + bytes) spread over all warps of the grid (lane 0 of each warp issues)
__device__ __forceinline__ void prefetch_region(const unsigned char* base, unsigned long long bytes, int gw, int n_gw) {
  if (base == nullptr || bytes == 0) return;
  const unsigned long long per = ((bytes / (unsigned long long)n_gw) + 127ull) & ~127ull;
  const unsigned long long lo = (unsigned long long)gw * per;
  if (lo >= bytes) return;
  const unsigned long long hi = lo + per < bytes ? lo + per : bytes;
  for (unsigned long long o = lo; o < hi; o += 4096ull) {
    const unsigned long long n = hi - o < 4096ull ? hi - o : 4096ull;
    bulk_prefetch_l2(base + o, (uint32_t)(n & ~15ull));
  }
}

__device__ __forceinline__ uint32_t float_key(float f) {  // monotone: larger float -> larger key
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

enum Mode { MODE_PLAIN = 0, MODE_GATED = 1, MODE_QKV = 2 };
static int g_l2_prefetch_mb = 0;  // see mdi_set_l2_prefetch_mb

// item -> the two weight rows a warp streams together
template <int MODE>
__device__ __forceinline__ void item_rows(const StreamArgs& a, int it, const bf16*& wa, const bf16*& wb) {
  if (MODE == MODE_GATED) {
    wa = a.W + (size_t)it * a.K;
    wb = a.W2 + (size_t)it * a.K;
  } else if (MODE == MODE_PLAIN) {
    wa = a.W + (size_t)(2 * it) * a.K;
    wb = a.W + (size_t)min(2 * it + 1, a.N - 1) * a.K;
  } else {
    const int hs = a.head_size, half_hs = hs / 2, ne = a.rope_n_elem, half_ne = ne / 2;
    const int j = it / half_hs, i = it % half_hs;
    int ra, rb;
    if (i < half_ne) { ra = i; rb = i + half_ne; }
    else { const int t = i - half_ne; ra = ne + 2 * t; rb = ra + 1; }
    wa = a.W + ((size_t)j * hs + ra) * a.K;
    wb = a.W + ((size_t)j * hs + rb) * a.K;
  }
}

// lane 0 of the owning warp: bias / activation / residual / RoPE / KV append, then the store
template <int MODE>
__device__ __forceinline__ void item_epilogue(const StreamArgs& a, int it, float da, float db, int slot, int pos,
                                              const bf16* res, unsigned int* hist_s, unsigned long long& best,
                                              const float* pre_res = nullptr) {
  if (MODE == MODE_GATED) {
    if (a.bias) da += __bfloat162float(a.bias[it]);
    if (a.bias2) db += __bfloat162float(a.bias2[it]);
    // the eager model rounds each projection to bf16, then act(a) (bf16) * b (bf16)
    const float fa = round_bf16(da), fb = round_bf16(db);
    float g;
    if (a.act == ACT_SILU_GATE) g = silu(fa);
    else if (a.act == ACT_GELU_TANH_GATE) g = gelu_tanh(fa);
    else g = gelu_erf(fa);
    const float out = round_bf16(g) * fb;
    if (a.out_fp32) reinterpret_cast<float*>(a.y)[(size_t)slot * a.y_slot_stride + it] = out;
    else reinterpret_cast<bf16*>(a.y)[(size_t)slot * a.y_slot_stride + it] = __float2bfloat16_rn(out);
  } else if (MODE == MODE_PLAIN) {
    const float o[2] = {da, db};
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int row = 2 * it + j;
      if (row >= a.N) continue;
      float v = o[j];
      if (a.bias) v += __bfloat162float(a.bias[row]);
      if (a.act == ACT_GELU_TANH) v = gelu_tanh(round_bf16(v));
      else if (a.act == ACT_GELU_ERF) v = gelu_erf(round_bf16(v));
      if (res) v = round_bf16(v) + (pre_res ? pre_res[j] : __bfloat162float(res[row]));
      // fp32 output = logits: bf16 values like nn.Linear would produce, kept in fp32 for the sampler
      if (a.out_fp32) {
        v = round_bf16(v);
        reinterpret_cast<float*>(a.y)[(size_t)slot * a.y_slot_stride + row] = v;
        if (hist_s) {  // sampling statistics: CTA-local histogram + per-warp arg-max, flushed once at the end
          const uint32_t key = float_key(v);
          atomicAdd(hist_s + (key >> 20), 1u);
          const unsigned long long packed = ((unsigned long long)key << 32) | (unsigned long long)(0xffffffffu - (uint32_t)row);
          best = packed > best ? packed : best;
        }
      } else {
        reinterpret_cast<bf16*>(a.y)[(size_t)slot * a.y_slot_stride + row] = __float2bfloat16_rn(v);
      }
    }
  } else {
    const int hs = a.head_size, half_hs = hs / 2, ne = a.rope_n_elem, half_ne = ne / 2;
    const int qpk = a.n_head / a.n_groups;
    const int j = it / half_hs, i = it % half_hs;
    int ra, rb;
    const bool rot_pair = i < half_ne;
    if (rot_pair) { ra = i; rb = i + half_ne; }
    else { const int t = i - half_ne; ra = ne + 2 * t; rb = ra + 1; }
    const size_t row0 = (size_t)j * hs;
    if (a.bias) { da += __bfloat162float(a.bias[row0 + ra]); db += __bfloat162float(a.bias[row0 + rb]); }
    da = round_bf16(da);
    db = round_bf16(db);
    const int g = j / (qpk + 2), s = j % (qpk + 2);
    if (rot_pair && s <= qpk) {  // q and k heads rotate, v does not
      const float ca = a.cos[(size_t)pos * ne + ra], sa = a.sin[(size_t)pos * ne + ra];
      const float cb = a.cos[(size_t)pos * ne + rb], sb = a.sin[(size_t)pos * ne + rb];
      const float na = da * ca - db * sa;
      const float nb = db * cb + da * sb;
      da = na; db = nb;
    }
    bf16* dst;
    if (s < qpk) dst = a.q_out + (size_t)(g * qpk + s) * hs;
    else {
      const size_t which = (s == qpk) ? 0 : 1;
      dst = a.kv + ((((size_t)slot * 2 + which) * a.n_groups + g) * a.max_seq + pos) * hs;
    }
    dst[ra] = __float2bfloat16_rn(da);
    dst[rb] = __float2bfloat16_rn(db);
  }
}

constexpr int STAT_BINS = 4096;

// zero the CTA-local histogram (placed after everything else in dynamic smem)
__device__ __forceinline__ unsigned int* stats_begin(const StreamArgs& a, unsigned char* smem_after) {
  if (a.hist == nullptr) return nullptr;
  unsigned int* h = reinterpret_cast<unsigned int*>(smem_after);
  for (int i = threadIdx.x; i < STAT_BINS; i += LIN_THREADS) h[i] = 0;
  return h;  // visibility: the callers __syncthreads() in stage_input before any epilogue runs
}
// one global atomic per populated bin per CTA, one atomicMax per warp
__device__ __forceinline__ void stats_flush(const StreamArgs& a, unsigned int* hist_s, unsigned long long best) {
  if (hist_s == nullptr) return;
  __syncthreads();
  for (int i = threadIdx.x; i < STAT_BINS; i += LIN_THREADS) {
    const unsigned int c = hist_s[i];
    if (c) atomicAdd(a.hist + i, c);
  }
  if ((threadIdx.x & 31) == 0 && best != 0ull && a.amax) atomicMax(a.amax, best);
}

// ---- variant A: register-streamed (LDG.128 batches) ---------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(LIN_THREADS, 3) stream_ldg_kernel(const StreamArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  bf16* xs = reinterpret_cast<bf16*>(smem_raw);
  __shared__ float red[LIN_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * LIN_WARPS + warp, n_gw = gridDim.x * LIN_WARPS;
  trace_mark(a.trace, 0, true);

  // pull this warp's first rows towards L2 while we (possibly) wait for the previous kernel / hop
  if (gw < a.n_items) {
    const bf16 *wa, *wb;
    item_rows<MODE>(a, gw, wa, wb);
    for (int off = lane * 64; off < a.K; off += 32 * 64) { prefetch_l2(wa + off); prefetch_l2(wb + off); }
  }
  unsigned int* hist_s = stats_begin(a, smem_raw + (size_t)((a.K + 63) / 64) * 64 * sizeof(bf16));
  unsigned long long best = 0ull;
  if (a.dep_wait.flag) dep_wait(a.dep_wait, a.ctx); else pdl_wait_prior();
  hop_wait(a.wait, a.ctx);
  trace_mark(a.trace, 1, true);
  const int slot = a.ctx ? a.ctx[MDI_CTX_SLOT] : 0, pos = a.ctx ? a.ctx[MDI_CTX_POS] : 0;
  stage_input(a.x + (size_t)slot * a.x_slot_stride, a.norm_w, a.eps, a.unit_offset, a.K, xs, red, a.norm_b, a.layer_norm);
  trace_mark(a.trace, 2, false);
  pdl_launch_dependents();

  const bf16* res = a.residual ? a.residual + (size_t)slot * a.res_slot_stride : nullptr;
  const uint4* xv = reinterpret_cast<const uint4*>(xs);
  const int nvec = a.K / 8;
  for (int it = gw; it < a.n_items; it += n_gw) {
    const bf16 *wa, *wb;
    item_rows<MODE>(a, it, wa, wb);
    float da, db;
    warp_dot2(wa, wb, xv, nvec, lane, da, db);
    if (lane == 0) item_epilogue<MODE>(a, it, da, db, slot, pos, res, hist_s, best);
  }
  stats_flush(a, hist_s, best);
  if (a.hop_row) hop_signal_copy(a.signal, a.ctx, reinterpret_cast<const bf16*>(a.y) + (size_t)slot * a.y_slot_stride,
                                 a.hop_row + (size_t)slot * a.hop_slot_stride, a.N,
                                 a.hop_pre ? a.hop_pre + (size_t)slot * a.hop_pre_slot_stride : nullptr, a.hop_pre_elems);
  else hop_signal(a.signal, a.ctx);
  dep_signal(a.dep_signal, a.ctx);
  trace_mark(a.trace, 3, true);
  trace_mark(a.trace, 4, false);
}

// ---- variant B: bulk-copy streamed (cp.async.bulk -> smem ring per warp, mbarrier completion) -------
// Every warp is its own producer: lane 0 keeps STAGES-1 row-chunk copies in flight through the TMA
// engine (no registers held by in-flight data), all lanes consume from shared memory.  The first
// copies are issued BEFORE the hop wait / input staging, so weight streaming overlaps both.
constexpr int TS_CHUNK = 1024;  // elements per row chunk (2 KB)

template <int MODE, int STAGES>
__global__ void __launch_bounds__(LIN_THREADS, STAGES == 2 ? 3 : (STAGES == 3 ? 2 : 1)) stream_bulk_kernel(const StreamArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * LIN_WARPS + warp, n_gw = gridDim.x * LIN_WARPS;
  trace_mark(a.trace, 0, true);
  // layout: ring [WARPS][STAGES][2][TS_CHUNK] bf16 | x [K] bf16 | mbar [WARPS][STAGES] u64
  bf16* ring = reinterpret_cast<bf16*>(smem_raw) + (size_t)warp * STAGES * 2 * TS_CHUNK;
  bf16* xs = reinterpret_cast<bf16*>(smem_raw) + (size_t)LIN_WARPS * STAGES * 2 * TS_CHUNK;
  uint64_t* bars = reinterpret_cast<uint64_t*>(xs + ((a.K + 63) / 64) * 64) + warp * STAGES;
  __shared__ float red[LIN_WARPS];

  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncwarp();

  const int n_chunks = (a.K + TS_CHUNK - 1) / TS_CHUNK;
  const int n_my = gw < a.n_items ? (a.n_items - gw + n_gw - 1) / n_gw : 0;
  const int total = n_my * n_chunks;
  auto issue = [&](int f) {
    const int ii = f / n_chunks, c = f - ii * n_chunks;
    const bf16 *wa, *wb;
    item_rows<MODE>(a, gw + ii * n_gw, wa, wb);
    const int k0 = c * TS_CHUNK;
    const uint32_t bytes = (uint32_t)min(TS_CHUNK, a.K - k0) * 2u;
    const int st = f % STAGES;
    bf16* dst = ring + (size_t)st * 2 * TS_CHUNK;
    mbar_expect_tx(&bars[st], 2 * bytes);
    bulk_g2s(dst, wa + k0, bytes, &bars[st]);
    bulk_g2s(dst + TS_CHUNK, wb + k0, bytes, &bars[st]);
  };
  // Fill the WHOLE ring before waiting for the input: the previous kernel's tail (and, for o_proj, the
  // attention kernel's latency chain) then overlaps with up to STAGES x 4 KB per warp of weight traffic.
  if (lane == 0) {
    for (int f = 0; f < min(STAGES, total); ++f) issue(f);
    // ... and ask the TMA engine to pull the next chunks into L2, so that the ring refills issued after
    // the wait hit L2 (lower latency, ~2x HBM bandwidth) instead of DRAM
    for (int f = STAGES; f < min(total, STAGES + a.l2_pf_chunks); ++f) {
      const int ii = f / n_chunks, c = f - ii * n_chunks;
      const bf16 *wa, *wb;
      item_rows<MODE>(a, gw + ii * n_gw, wa, wb);
      const int k0 = c * TS_CHUNK;
      const uint32_t bytes = (uint32_t)min(TS_CHUNK, a.K - k0) * 2u;
      bulk_prefetch_l2(wa + k0, bytes);
      bulk_prefetch_l2(wb + k0, bytes);
    }
    prefetch_region(a.pf_a, a.pf_bytes, gw, n_gw);
    prefetch_region(a.pf_b, a.pf_bytes, gw, n_gw);
  }
  unsigned int* hist_s = stats_begin(a, reinterpret_cast<unsigned char*>(bars - warp * STAGES + LIN_WARPS * STAGES));
  unsigned long long best = 0ull;

  int slot = 0, pos = 0;
  if (a.ctx_early && a.ctx) { slot = a.ctx[MDI_CTX_SLOT]; pos = a.ctx[MDI_CTX_POS]; }  // off the post-wait chain
  uint4 w_pre[4];
  const bool have_w = a.norm_w != nullptr && !a.layer_norm && a.K <= STAGE_VPT * LIN_THREADS * 8;
  if (have_w) preload_norm_w(a.norm_w, a.K, w_pre);
  if (a.dep_wait.flag) dep_wait(a.dep_wait, a.ctx); else pdl_wait_prior();
  hop_wait(a.wait, a.ctx);
  trace_mark(a.trace, 1, true);
  if (!a.ctx_early && a.ctx) { slot = a.ctx[MDI_CTX_SLOT]; pos = a.ctx[MDI_CTX_POS]; }
  stage_input(a.x + (size_t)slot * a.x_slot_stride, a.norm_w, a.eps, a.unit_offset, a.K, xs, red, a.norm_b, a.layer_norm,
              have_w ? w_pre : nullptr);
  trace_mark(a.trace, 2, false);
  pdl_launch_dependents();

  const bf16* res = a.residual ? a.residual + (size_t)slot * a.res_slot_stride : nullptr;
  float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
  int c = 0, ii = 0;
  float pre_res[2] = {0.f, 0.f};  // the item's two residual values, requested when the item starts (off the tail)
  for (int f = 0; f < total; ++f) {
    const int st = f % STAGES;
    if (MODE == MODE_PLAIN && c == 0 && res != nullptr && lane == 0) {
      const int row = 2 * (gw + ii * n_gw);
      pre_res[0] = __bfloat162float(res[row]);
      pre_res[1] = row + 1 < a.N ? __bfloat162float(res[row + 1]) : 0.f;
    }
    mbar_wait(&bars[st], (uint32_t)((f / STAGES) & 1));
    const int k0 = c * TS_CHUNK;
    const int nv = min(TS_CHUNK, a.K - k0) / 8;
    const uint4* wa = reinterpret_cast<const uint4*>(ring + (size_t)st * 2 * TS_CHUNK);
    const uint4* wb = wa + TS_CHUNK / 8;
    const uint4* xv = reinterpret_cast<const uint4*>(xs + k0);
    if (nv == TS_CHUNK / 8) {
#pragma unroll
      for (int u = 0; u < TS_CHUNK / 8 / 32; ++u) {
        const uint4 x = xv[lane + 32 * u];
        if (u & 1) { a1 += dot8(wa[lane + 32 * u], x); b1 += dot8(wb[lane + 32 * u], x); }
        else { a0 += dot8(wa[lane + 32 * u], x); b0 += dot8(wb[lane + 32 * u], x); }
      }
    } else {
      for (int v = lane; v < nv; v += 32) { const uint4 x = xv[v]; a0 += dot8(wa[v], x); b0 += dot8(wb[v], x); }
    }
    if (++c == n_chunks) {
      const float da = warp_sum(a0 + a1), db = warp_sum(b0 + b1);
      if (lane == 0) item_epilogue<MODE>(a, gw + ii * n_gw, da, db, slot, pos, res, hist_s, best,
                                         (MODE == MODE_PLAIN && res != nullptr) ? pre_res : nullptr);
      a0 = a1 = b0 = b1 = 0.f;
      c = 0;
      ++ii;
    }
    __syncwarp();  // everyone is done with stage `st` ...
    if (lane == 0 && f + STAGES < total) issue(f + STAGES);  // ... so it can be refilled right away
  }
  stats_flush(a, hist_s, best);
  if (a.hop_row) hop_signal_copy(a.signal, a.ctx, reinterpret_cast<const bf16*>(a.y) + (size_t)slot * a.y_slot_stride,
                                 a.hop_row + (size_t)slot * a.hop_slot_stride, a.N,
                                 a.hop_pre ? a.hop_pre + (size_t)slot * a.hop_pre_slot_stride : nullptr, a.hop_pre_elems);
  else hop_signal(a.signal, a.ctx);
  dep_signal(a.dep_signal, a.ctx);
  trace_mark(a.trace, 3, true);
  trace_mark(a.trace, 4, false);
}

// ---- fp8 (e4m3) block-scaled weights, bf16/fp16 activations ("W8A16") ------------------------------
// Decode is bandwidth-bound, so serving in fp8 halves the bytes per token.  Weights are dequantised
// in registers: cvt.e4m3x2 -> f16x2, HFMA2 against the fp16 activation vector over one 16-element
// vector, then one fp32 FMA with the block scale (block = 128 elements of K, so a lane's 16-byte
// vector never straddles two scales).
template <int MODE>
__device__ __forceinline__ void item_rows_fp8(const StreamArgs& a, int it, const unsigned char*& wa, const unsigned char*& wb,
                                              const float*& sa, const float*& sb) {
  const unsigned char* W = reinterpret_cast<const unsigned char*>(a.W);
  const int nblk = a.K / 128;
  size_t ra, rb;
  const unsigned char* W2 = W;
  const float* S2 = a.wscale;
  if (MODE == MODE_GATED) {
    ra = rb = (size_t)it;
    W2 = reinterpret_cast<const unsigned char*>(a.W2);
    S2 = a.wscale2;
  } else if (MODE == MODE_PLAIN) {
    ra = (size_t)2 * it;
    rb = (size_t)min(2 * it + 1, a.N - 1);
  } else {
    const int hs = a.head_size, half_hs = hs / 2, ne = a.rope_n_elem, half_ne = ne / 2;
    const int j = it / half_hs, i = it % half_hs;
    int r0, r1;
    if (i < half_ne) { r0 = i; r1 = i + half_ne; }
    else { const int t = i - half_ne; r0 = ne + 2 * t; r1 = r0 + 1; }
    ra = (size_t)j * hs + r0;
    rb = (size_t)j * hs + r1;
  }
  wa = W + ra * a.K;
  wb = W2 + rb * a.K;
  sa = a.wscale + ra * nblk;
  sb = S2 + rb * nblk;
}

__device__ __forceinline__ float dot16_fp8(const uint4& w, const uint4& x0, const uint4& x1) {
  const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
  const uint32_t xx[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
  __half2 acc = __float2half2_rn(0.f);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2_raw lo = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(ww[i] & 0xffffu), __NV_E4M3);
    const __half2_raw hi = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)(ww[i] >> 16), __NV_E4M3);
    acc = __hfma2(*reinterpret_cast<const __half2*>(&lo), *reinterpret_cast<const __half2*>(&xx[2 * i]), acc);
    acc = __hfma2(*reinterpret_cast<const __half2*>(&hi), *reinterpret_cast<const __half2*>(&xx[2 * i + 1]), acc);
  }
  const float2 f = __half22float2(acc);
  return f.x + f.y;
}

__device__ __forceinline__ void warp_dot2_fp8(const unsigned char* wa, const unsigned char* wb, const float* sa, const float* sb,
                                              const uint4* __restrict__ xh, int nvec, int lane, float& out_a, float& out_b) {
  const uint4* pa = reinterpret_cast<const uint4*>(wa);
  const uint4* pb = reinterpret_cast<const uint4*>(wb);
  float a0 = 0.f, b0 = 0.f;
  int v = lane;  // 16 fp8 weights per vector
  for (; v + 32 * (LIN_UNROLL - 1) < nvec; v += 32 * LIN_UNROLL) {
    uint4 ra[LIN_UNROLL], rb[LIN_UNROLL];
    float fa[LIN_UNROLL], fb[LIN_UNROLL];
#pragma unroll
    for (int u = 0; u < LIN_UNROLL; ++u) {
      ra[u] = ldg_stream(pa + v + 32 * u);
      rb[u] = ldg_stream(pb + v + 32 * u);
      fa[u] = __ldg(sa + ((v + 32 * u) >> 3));
      fb[u] = __ldg(sb + ((v + 32 * u) >> 3));
    }
#pragma unroll
    for (int u = 0; u < LIN_UNROLL; ++u) {
      const uint4 x0 = xh[2 * (v + 32 * u)], x1 = xh[2 * (v + 32 * u) + 1];
      a0 = fmaf(fa[u], dot16_fp8(ra[u], x0, x1), a0);
      b0 = fmaf(fb[u], dot16_fp8(rb[u], x0, x1), b0);
    }
  }
  for (; v < nvec; v += 32) {
    const uint4 x0 = xh[2 * v], x1 = xh[2 * v + 1];
    a0 = fmaf(__ldg(sa + (v >> 3)), dot16_fp8(ldg_stream(pa + v), x0, x1), a0);
    b0 = fmaf(__ldg(sb + (v >> 3)), dot16_fp8(ldg_stream(pb + v), x0, x1), b0);
  }
  out_a = warp_sum(a0);
  out_b = warp_sum(b0);
}

template <int MODE>
__global__ void __launch_bounds__(LIN_THREADS, 3) stream_ldg_fp8_kernel(const StreamArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  bf16* xs = reinterpret_cast<bf16*>(smem_raw);
  __shared__ float red[LIN_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * LIN_WARPS + warp, n_gw = gridDim.x * LIN_WARPS;
  trace_mark(a.trace, 0, true);
  if (gw < a.n_items) {
    const unsigned char *wa, *wb;
    const float *sa, *sb;
    item_rows_fp8<MODE>(a, gw, wa, wb, sa, sb);
    for (int off = lane * 128; off < a.K; off += 32 * 128) { prefetch_l2(wa + off); prefetch_l2(wb + off); }
  }
  unsigned int* hist_s = stats_begin(a, smem_raw + (size_t)((a.K + 63) / 64) * 64 * sizeof(bf16));
  unsigned long long best = 0ull;
  if (a.dep_wait.flag) dep_wait(a.dep_wait, a.ctx); else pdl_wait_prior();
  hop_wait(a.wait, a.ctx);
  trace_mark(a.trace, 1, true);
  const int slot = a.ctx ? a.ctx[MDI_CTX_SLOT] : 0, pos = a.ctx ? a.ctx[MDI_CTX_POS] : 0;
  stage_input(a.x + (size_t)slot * a.x_slot_stride, a.norm_w, a.eps, a.unit_offset, a.K, xs, red, a.norm_b, a.layer_norm);
  // activations to fp16 in place (same 2 bytes) for the HFMA2 inner product
  __half* xh16 = reinterpret_cast<__half*>(smem_raw);
  for (int i = threadIdx.x; i < a.K; i += LIN_THREADS) xh16[i] = __float2half_rn(__bfloat162float(xs[i]));
  __syncthreads();
  trace_mark(a.trace, 2, false);
  pdl_launch_dependents();

  const bf16* res = a.residual ? a.residual + (size_t)slot * a.res_slot_stride : nullptr;
  const uint4* xv = reinterpret_cast<const uint4*>(smem_raw);
  const int nvec = a.K / 16;
  for (int it = gw; it < a.n_items; it += n_gw) {
    const unsigned char *wa, *wb;
    const float *sa, *sb;
    item_rows_fp8<MODE>(a, it, wa, wb, sa, sb);
    float da, db;
    warp_dot2_fp8(wa, wb, sa, sb, xv, nvec, lane, da, db);
    if (lane == 0) item_epilogue<MODE>(a, it, da, db, slot, pos, res, hist_s, best);
  }
  stats_flush(a, hist_s, best);
  if (a.hop_row) hop_signal_copy(a.signal, a.ctx, reinterpret_cast<const bf16*>(a.y) + (size_t)slot * a.y_slot_stride,
                                 a.hop_row + (size_t)slot * a.hop_slot_stride, a.N,
                                 a.hop_pre ? a.hop_pre + (size_t)slot * a.hop_pre_slot_stride : nullptr, a.hop_pre_elems);
  else hop_signal(a.signal, a.ctx);
  dep_signal(a.dep_signal, a.ctx);
  trace_mark(a.trace, 3, true);
  trace_mark(a.trace, 4, false);
}

// ---- fp8 weights through the bulk-copy (TMA engine) ring ------------------------------------------------
// Same per-warp mbarrier ring as the bf16 kernel, but a chunk is 2048 one-byte weights per row (so a stage
// holds the same 2 x 2 KB), the activations sit in shared memory as fp16 and the block scales of the
// chunk are requested before the mbarrier wait so their L2 latency hides behind the copy.
constexpr int TS8 = 2048;
template <int MODE, int STAGES>
__global__ void __launch_bounds__(LIN_THREADS, 3) stream_bulk_fp8_kernel(const StreamArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * LIN_WARPS + warp, n_gw = gridDim.x * LIN_WARPS;
  trace_mark(a.trace, 0, true);
  // layout: ring [WARPS][STAGES][2][TS8] bytes | x [K] fp16 | mbar [WARPS][STAGES] u64
  unsigned char* ring = smem_raw + (size_t)warp * STAGES * 2 * TS8;
  bf16* xs = reinterpret_cast<bf16*>(smem_raw + (size_t)LIN_WARPS * STAGES * 2 * TS8);
  uint64_t* bars = reinterpret_cast<uint64_t*>(xs + ((a.K + 63) / 64) * 64) + warp * STAGES;
  __shared__ float red[LIN_WARPS];
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncwarp();
  const int n_chunks = (a.K + TS8 - 1) / TS8;
  const int n_my = gw < a.n_items ? (a.n_items - gw + n_gw - 1) / n_gw : 0;
  const int total = n_my * n_chunks;
  auto issue = [&](int f) {
    const int ii = f / n_chunks, c = f - ii * n_chunks;
    const unsigned char *wa, *wb;
    const float *sa, *sb;
    item_rows_fp8<MODE>(a, gw + ii * n_gw, wa, wb, sa, sb);
    const int k0 = c * TS8;
    const uint32_t bytes = (uint32_t)min(TS8, a.K - k0);
    const int st = f % STAGES;
    unsigned char* dst = ring + (size_t)st * 2 * TS8;
    mbar_expect_tx(&bars[st], 2 * bytes);
    bulk_g2s(dst, wa + k0, bytes, &bars[st]);
    bulk_g2s(dst + TS8, wb + k0, bytes, &bars[st]);
  };
  if (lane == 0)
    for (int f = 0; f < min(STAGES, total); ++f) issue(f);
  unsigned int* hist_s = stats_begin(a, reinterpret_cast<unsigned char*>(bars - warp * STAGES + LIN_WARPS * STAGES));
  unsigned long long best = 0ull;

  int slot = 0, pos = 0;
  if (a.ctx_early && a.ctx) { slot = a.ctx[MDI_CTX_SLOT]; pos = a.ctx[MDI_CTX_POS]; }
  if (a.dep_wait.flag) dep_wait(a.dep_wait, a.ctx); else pdl_wait_prior();
  hop_wait(a.wait, a.ctx);
  trace_mark(a.trace, 1, true);
  if (!a.ctx_early && a.ctx) { slot = a.ctx[MDI_CTX_SLOT]; pos = a.ctx[MDI_CTX_POS]; }
  stage_input(a.x + (size_t)slot * a.x_slot_stride, a.norm_w, a.eps, a.unit_offset, a.K, xs, red, a.norm_b, a.layer_norm);
  __half* xh16 = reinterpret_cast<__half*>(xs);  // activations to fp16 in place for the HFMA2 inner product
  for (int i = threadIdx.x; i < a.K; i += LIN_THREADS) xh16[i] = __float2half_rn(__bfloat162float(xs[i]));
  __syncthreads();
  trace_mark(a.trace, 2, false);
  pdl_launch_dependents();

  const bf16* res = a.residual ? a.residual + (size_t)slot * a.res_slot_stride : nullptr;
  const uint4* xv = reinterpret_cast<const uint4*>(xs);
  float acc_a = 0.f, acc_b = 0.f;
  int c = 0, ii = 0;
  const float *sa = nullptr, *sb = nullptr;
  for (int f = 0; f < total; ++f) {
    const int st = f % STAGES;
    if (c == 0) {
      const unsigned char *wa, *wb;
      item_rows_fp8<MODE>(a, gw + ii * n_gw, wa, wb, sa, sb);
    }
    const int k0 = c * TS8;
    const int nv = min(TS8, a.K - k0) / 16;  // 16-weight vectors in this chunk (<= 128: 4 per lane)
    float fa[4], fb[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int v = lane + 32 * u;
      const int blk = (k0 >> 7) + (v >> 3);
      fa[u] = v < nv ? __ldg(sa + blk) : 0.f;
      fb[u] = v < nv ? __ldg(sb + blk) : 0.f;
    }
    mbar_wait(&bars[st], (uint32_t)((f / STAGES) & 1));
    const uint4* wa4 = reinterpret_cast<const uint4*>(ring + (size_t)st * 2 * TS8);
    const uint4* wb4 = wa4 + TS8 / 16;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int v = lane + 32 * u;
      if (v < nv) {
        const uint4 x0 = xv[(k0 >> 3) + 2 * v], x1 = xv[(k0 >> 3) + 2 * v + 1];
        acc_a = fmaf(fa[u], dot16_fp8(wa4[v], x0, x1), acc_a);
        acc_b = fmaf(fb[u], dot16_fp8(wb4[v], x0, x1), acc_b);
      }
    }
    if (++c == n_chunks) {
      const float da = warp_sum(acc_a), db = warp_sum(acc_b);
      if (lane == 0) item_epilogue<MODE>(a, gw + ii * n_gw, da, db, slot, pos, res, hist_s, best);
      acc_a = acc_b = 0.f;
      c = 0;
      ++ii;
    }
    __syncwarp();
    if (lane == 0 && f + STAGES < total) issue(f + STAGES);
  }
  stats_flush(a, hist_s, best);
  if (a.hop_row) hop_signal_copy(a.signal, a.ctx, reinterpret_cast<const bf16*>(a.y) + (size_t)slot * a.y_slot_stride,
                                 a.hop_row + (size_t)slot * a.hop_slot_stride, a.N,
                                 a.hop_pre ? a.hop_pre + (size_t)slot * a.hop_pre_slot_stride : nullptr, a.hop_pre_elems);
  else hop_signal(a.signal, a.ctx);
  dep_signal(a.dep_signal, a.ctx);
  trace_mark(a.trace, 3, true);
  trace_mark(a.trace, 4, false);
}

// ---- mixture of experts (LLaMAMoE, model.py:823-853) ------------------------------------------------------
// The reference routes with torch.topk + torch.where (a host sync per layer) and loops over ALL experts.  Here the
// router is one small kernel that leaves the token's expert ids and routing weights in device memory, and each of
// the token's `top` experts is one gate/up + one down launch whose weight base pointers are looked up through a
// device pointer table AFTER the dependency wait — so a decode step streams only the chosen experts' matrices
// (Mixtral: 2 of 8) and stays graph-capturable.  Rounding points follow the eager module: logits, routing weights,
// each expert's output and the running sum are bf16.
struct RouterArgs {
  const bf16* Wg;        // [E, K] router matrix
  const bf16* x;         // residual stream row: x + slot * x_slot_stride
  const bf16* norm_w;    // the block's norm_2
  const bf16* norm_b;
  int layer_norm;
  const int* ctx;
  long long x_slot_stride;
  int K, E, top;
  float eps;
  int unit_offset;
  int* sel;              // out [top]: chosen expert ids, best first
  float* wts;            // out [top]: softmax over the chosen logits (bf16 values)
  HopWait wait;
  unsigned long long* trace;
};

constexpr int MOE_MAX_EXPERTS = 256;
constexpr int MOE_MAX_TOP = 8;

__global__ void __launch_bounds__(LIN_THREADS) moe_router_kernel(const RouterArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  bf16* xs = reinterpret_cast<bf16*>(smem_raw);
  __shared__ float red[LIN_WARPS];
  __shared__ float logit_s[MOE_MAX_EXPERTS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  trace_mark(a.trace, 0, true);
  pdl_wait_prior();
  hop_wait(a.wait, a.ctx);
  trace_mark(a.trace, 1, true);
  const int slot = a.ctx ? a.ctx[MDI_CTX_SLOT] : 0;
  stage_input(a.x + (size_t)slot * a.x_slot_stride, a.norm_w, a.eps, a.unit_offset, a.K, xs, red, a.norm_b, a.layer_norm);
  trace_mark(a.trace, 2, false);
  pdl_launch_dependents();  // the expert kernels still wait for this grid to COMPLETE before they read sel / wts
  const uint4* xv = reinterpret_cast<const uint4*>(xs);
  const int nvec = a.K / 8;
  for (int e = warp; e < a.E; e += LIN_WARPS) {
    const uint4* w = reinterpret_cast<const uint4*>(a.Wg + (size_t)e * a.K);
    float s = 0.f;
    for (int v = lane; v < nvec; v += 32) s += dot8(__ldg(w + v), xv[v]);
    s = warp_sum(s);
    if (lane == 0) logit_s[e] = round_bf16(s);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int chosen[MOE_MAX_TOP];
    float val[MOE_MAX_TOP];
    for (int k = 0; k < a.top; ++k) {  // E is small: selection by repeated arg-max, lowest index wins ties
      int best = -1;
      float bv = 0.f;
      for (int e = 0; e < a.E; ++e) {
        bool taken = false;
        for (int j = 0; j < k; ++j) taken |= chosen[j] == e;
        if (!taken && (best < 0 || logit_s[e] > bv)) { best = e; bv = logit_s[e]; }
      }
      chosen[k] = best;
      val[k] = bv;
    }
    float den = 0.f;
    const float vmax = val[0];  // best first
    for (int k = 0; k < a.top; ++k) { val[k] = expf(val[k] - vmax); den += val[k]; }
    for (int k = 0; k < a.top; ++k) {
      a.sel[k] = chosen[k];
      a.wts[k] = round_bf16(val[k] / den);
    }
  }
  trace_mark(a.trace, 3, true);
  trace_mark(a.trace, 4, false);
}

struct MoeArgs {
  const bf16* const* w;   // [E] device pointers: fc_1 (gated pass) or proj (down pass) of every expert
  const bf16* const* w2;  // [E] fc_2 of every expert (gated pass) or null
  const int* sel;         // the router's output
  const float* wts;
  int k;                  // which of the token's experts this launch computes
  const bf16* prev;       // down pass: running bf16 sum of the previous experts' weighted outputs [N], or null
  int sel_early;          // the router finished >= 2 launches ago: sel / wts may be read BEFORE the dependency wait
};

__device__ __forceinline__ void moe_down_store(const StreamArgs& a, int row, int slot, float acc, float route, bool has_prev,
                                               float prev, bool has_res, float res) {
  float v = round_bf16(round_bf16(acc) * route);
  if (has_prev) v = round_bf16(v + prev);
  if (has_res) v += res;
  reinterpret_cast<bf16*>(a.y)[(size_t)slot * a.y_slot_stride + row] = __float2bfloat16_rn(v);
}

// Register-streamed like stream_ldg_kernel, but nothing of the weights may be touched before the wait: which
// matrices to read is the router's output.  GATED: h = act(fc_1 x) * fc_2 x of expert sel[k].  PLAIN (down):
// y = bf16(wts[k] * bf16(proj h)) (+ prev) (+ residual on the token's last expert).
template <int MODE>
__global__ void __launch_bounds__(LIN_THREADS, 3) moe_stream_kernel(const StreamArgs a, const MoeArgs m) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  bf16* xs = reinterpret_cast<bf16*>(smem_raw);
  __shared__ float red[LIN_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * LIN_WARPS + warp, n_gw = gridDim.x * LIN_WARPS;
  trace_mark(a.trace, 0, true);
  unsigned long long best = 0ull;
  pdl_wait_prior();
  trace_mark(a.trace, 1, true);
  const int slot = a.ctx ? a.ctx[MDI_CTX_SLOT] : 0, pos = a.ctx ? a.ctx[MDI_CTX_POS] : 0;
  const int e = m.sel[m.k];
  const float route = m.wts[m.k];
  const bf16* W = m.w[e];
  const bf16* W2 = MODE == MODE_GATED ? m.w2[e] : nullptr;
  stage_input(a.x + (size_t)slot * a.x_slot_stride, a.norm_w, a.eps, a.unit_offset, a.K, xs, red, a.norm_b, a.layer_norm);
  trace_mark(a.trace, 2, false);
  pdl_launch_dependents();

  const bf16* res = a.residual ? a.residual + (size_t)slot * a.res_slot_stride : nullptr;
  const uint4* xv = reinterpret_cast<const uint4*>(xs);
  const int nvec = a.K / 8;
  for (int it = gw; it < a.n_items; it += n_gw) {
    const bf16 *wa, *wb;
    if (MODE == MODE_GATED) {
      wa = W + (size_t)it * a.K;
      wb = W2 + (size_t)it * a.K;
    } else {
      wa = W + (size_t)(2 * it) * a.K;
      wb = W + (size_t)min(2 * it + 1, a.N - 1) * a.K;
    }
    float da, db;
    warp_dot2(wa, wb, xv, nvec, lane, da, db);
    if (lane != 0) continue;
    if (MODE == MODE_GATED) {
      item_epilogue<MODE_GATED>(a, it, da, db, slot, pos, res, nullptr, best);
    } else {
      const float o[2] = {da, db};
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int row = 2 * it + j;
        if (row >= a.N) continue;
        moe_down_store(a, row, slot, o[j], route, m.prev != nullptr, m.prev ? __bfloat162float(m.prev[row]) : 0.f,
                       res != nullptr, res ? __bfloat162float(res[row]) : 0.f);
      }
    }
  }
  if (a.hop_row) hop_signal_copy(a.signal, a.ctx, reinterpret_cast<const bf16*>(a.y) + (size_t)slot * a.y_slot_stride,
                                 a.hop_row + (size_t)slot * a.hop_slot_stride, a.N, nullptr, 0);
  else hop_signal(a.signal, a.ctx);
  trace_mark(a.trace, 3, true);
  trace_mark(a.trace, 4, false);
}

// The same passes through the per-warp bulk-copy ring of stream_bulk_kernel.  Every launch but the first one after
// the router (`sel_early`) knows its expert before the dependency wait and fills its ring while the previous kernel
// drains, exactly like the dense kernels; the first one resolves its pointers right after the wait.
template <int MODE, int STAGES>
__global__ void __launch_bounds__(LIN_THREADS, STAGES == 2 ? 3 : 2) moe_bulk_kernel(const StreamArgs a, const MoeArgs m) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * LIN_WARPS + warp, n_gw = gridDim.x * LIN_WARPS;
  trace_mark(a.trace, 0, true);
  bf16* ring = reinterpret_cast<bf16*>(smem_raw) + (size_t)warp * STAGES * 2 * TS_CHUNK;
  bf16* xs = reinterpret_cast<bf16*>(smem_raw) + (size_t)LIN_WARPS * STAGES * 2 * TS_CHUNK;
  uint64_t* bars = reinterpret_cast<uint64_t*>(xs + ((a.K + 63) / 64) * 64) + warp * STAGES;
  __shared__ float red[LIN_WARPS];
  if (lane == 0) {
#pragma unroll
    for (int s = 0; s < STAGES; ++s) mbar_init(&bars[s], 1);
    mbar_fence_init();
  }
  __syncwarp();

  const int n_chunks = (a.K + TS_CHUNK - 1) / TS_CHUNK;
  const int n_my = gw < a.n_items ? (a.n_items - gw + n_gw - 1) / n_gw : 0;
  const int total = n_my * n_chunks;
  const bf16 *W = nullptr, *W2 = nullptr;
  float route = 0.f;
  auto resolve = [&]() {
    const int e = m.sel[m.k];
    route = m.wts[m.k];
    W = m.w[e];
    if (MODE == MODE_GATED) W2 = m.w2[e];
  };
  auto issue = [&](int f) {
    const int ii = f / n_chunks, c = f - ii * n_chunks;
    const int it = gw + ii * n_gw;
    const bf16 *wa, *wb;
    if (MODE == MODE_GATED) {
      wa = W + (size_t)it * a.K;
      wb = W2 + (size_t)it * a.K;
    } else {
      wa = W + (size_t)(2 * it) * a.K;
      wb = W + (size_t)min(2 * it + 1, a.N - 1) * a.K;
    }
    const int k0 = c * TS_CHUNK;
    const uint32_t bytes = (uint32_t)min(TS_CHUNK, a.K - k0) * 2u;
    const int st = f % STAGES;
    bf16* dst = ring + (size_t)st * 2 * TS_CHUNK;
    mbar_expect_tx(&bars[st], 2 * bytes);
    bulk_g2s(dst, wa + k0, bytes, &bars[st]);
    bulk_g2s(dst + TS_CHUNK, wb + k0, bytes, &bars[st]);
  };
  if (m.sel_early) {
    resolve();
    if (lane == 0)
      for (int f = 0; f < min(STAGES, total); ++f) issue(f);
  }
  uint4 w_pre[4];
  const bool have_w = a.norm_w != nullptr && !a.layer_norm && a.K <= STAGE_VPT * LIN_THREADS * 8;
  if (have_w) preload_norm_w(a.norm_w, a.K, w_pre);
  pdl_wait_prior();
  trace_mark(a.trace, 1, true);
  if (!m.sel_early) {
    resolve();
    if (lane == 0)
      for (int f = 0; f < min(STAGES, total); ++f) issue(f);
  }
  const int slot = a.ctx ? a.ctx[MDI_CTX_SLOT] : 0, pos = a.ctx ? a.ctx[MDI_CTX_POS] : 0;
  stage_input(a.x + (size_t)slot * a.x_slot_stride, a.norm_w, a.eps, a.unit_offset, a.K, xs, red, a.norm_b, a.layer_norm,
              have_w ? w_pre : nullptr);
  trace_mark(a.trace, 2, false);
  pdl_launch_dependents();

  const bf16* res = a.residual ? a.residual + (size_t)slot * a.res_slot_stride : nullptr;
  float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
  int c = 0, ii = 0;
  float pre_res[2] = {0.f, 0.f}, pre_prev[2] = {0.f, 0.f};  // requested when the item starts (off the tail)
  unsigned long long best = 0ull;
  for (int f = 0; f < total; ++f) {
    const int st = f % STAGES;
    if (MODE == MODE_PLAIN && c == 0 && lane == 0) {
      const int row = 2 * (gw + ii * n_gw);
      const bool two = row + 1 < a.N;
      if (res != nullptr) { pre_res[0] = __bfloat162float(res[row]); pre_res[1] = two ? __bfloat162float(res[row + 1]) : 0.f; }
      if (m.prev != nullptr) { pre_prev[0] = __bfloat162float(m.prev[row]); pre_prev[1] = two ? __bfloat162float(m.prev[row + 1]) : 0.f; }
    }
    mbar_wait(&bars[st], (uint32_t)((f / STAGES) & 1));
    const int k0 = c * TS_CHUNK;
    const int nv = min(TS_CHUNK, a.K - k0) / 8;
    const uint4* wa = reinterpret_cast<const uint4*>(ring + (size_t)st * 2 * TS_CHUNK);
    const uint4* wb = wa + TS_CHUNK / 8;
    const uint4* xv = reinterpret_cast<const uint4*>(xs + k0);
    if (nv == TS_CHUNK / 8) {
#pragma unroll
      for (int u = 0; u < TS_CHUNK / 8 / 32; ++u) {
        const uint4 x = xv[lane + 32 * u];
        if (u & 1) { a1 += dot8(wa[lane + 32 * u], x); b1 += dot8(wb[lane + 32 * u], x); }
        else { a0 += dot8(wa[lane + 32 * u], x); b0 += dot8(wb[lane + 32 * u], x); }
      }
    } else {
      for (int v = lane; v < nv; v += 32) { const uint4 x = xv[v]; a0 += dot8(wa[v], x); b0 += dot8(wb[v], x); }
    }
    if (++c == n_chunks) {
      const float da = warp_sum(a0 + a1), db = warp_sum(b0 + b1);
      if (lane == 0) {
        const int it = gw + ii * n_gw;
        if (MODE == MODE_GATED) {
          item_epilogue<MODE_GATED>(a, it, da, db, slot, pos, res, nullptr, best);
        } else {
          moe_down_store(a, 2 * it, slot, da, route, m.prev != nullptr, pre_prev[0], res != nullptr, pre_res[0]);
          if (2 * it + 1 < a.N)
            moe_down_store(a, 2 * it + 1, slot, db, route, m.prev != nullptr, pre_prev[1], res != nullptr, pre_res[1]);
        }
      }
      a0 = a1 = b0 = b1 = 0.f;
      c = 0;
      ++ii;
    }
    __syncwarp();
    if (lane == 0 && f + STAGES < total) issue(f + STAGES);
  }
  if (a.hop_row) hop_signal_copy(a.signal, a.ctx, reinterpret_cast<const bf16*>(a.y) + (size_t)slot * a.y_slot_stride,
                                 a.hop_row + (size_t)slot * a.hop_slot_stride, a.N, nullptr, 0);
  else hop_signal(a.signal, a.ctx);
  trace_mark(a.trace, 3, true);
  trace_mark(a.trace, 4, false);
}

// 0 = register-streamed (default), 2 = bulk-copy ring x2 (mdi_set_moe_variant).  Measured on Mixtral-8x7B shapes the two are
// within run-to-run noise (302.9 vs 309.4 us per 2-layer step, profiles/r2/moe_decode_bench_v*.json), so the simpler one stays.
static int g_moe_variant = 0;

static int g_num_sms = 0;
static int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

template <typename Kern>
static int launch_pdl(Kern kern, const StreamArgs& args, int grid, size_t smem, cudaStream_t stream, int use_pdl) {
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(LIN_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  return (int)cudaLaunchKernelEx(&cfg, kern, args);
}

// variant: 0 = LDG register streaming (grid = sms * ctas_per_sm), 1 = bulk-copy ring with 4 stages
// (1 CTA/SM), 2 / 3 = bulk-copy ring with 2 / 3 stages (grid = sms * as many CTAs as shared memory allows,
// capped by ctas_per_sm).
template <int MODE>
static int launch_stream(const StreamArgs& a, int variant, int ctas_per_sm, int use_pdl, cudaStream_t stream) {
  const int sms = num_sms();
  const int grid_override = ctas_per_sm < 0 ? -ctas_per_sm : 0;  // experiments: ctas_per_sm = -G forces a grid of G CTAs
  if (ctas_per_sm <= 0) ctas_per_sm = 4;
  const int max_useful = (a.n_items + LIN_WARPS - 1) / LIN_WARPS;
  if (a.wscale != nullptr) {  // fp8 block-scaled weights
    if (a.K % 128 != 0) return -2;
    if (variant != 0) {  // bulk-copy ring (2 stages), like the bf16 default
      const size_t smem8b = (size_t)LIN_WARPS * 2 * 2 * TS8 + (size_t)((a.K + 63) / 64) * 64 * 2 + (size_t)LIN_WARPS * 2 * 8 +
                            (a.hist ? STAT_BINS * 4 : 0);
      if (smem8b <= 227 * 1024) {
        const int per_sm8 = max(1, min(ctas_per_sm, (int)((227 * 1024) / (smem8b + 1024))));
        const int grid8 = grid_override > 0 ? min(grid_override, max_useful) : max(1, min(sms * per_sm8, max_useful));
        return launch_pdl(stream_bulk_fp8_kernel<MODE, 2>, a, grid8, smem8b, stream, use_pdl);
      }
    }
    const int grid = max(1, min(sms * min(ctas_per_sm, 3), max_useful));
    const size_t smem8 = (size_t)((a.K + 63) / 64) * 64 * sizeof(bf16) + (a.hist ? STAT_BINS * 4 : 0);
    return launch_pdl(stream_ldg_fp8_kernel<MODE>, a, grid, smem8, stream, use_pdl);
  }
  if (variant == 0) {
    const int grid = max(1, min(sms * ctas_per_sm, max_useful));
    const size_t smem0 = (size_t)((a.K + 63) / 64) * 64 * sizeof(bf16) + (a.hist ? STAT_BINS * 4 : 0);
    return launch_pdl(stream_ldg_kernel<MODE>, a, grid, smem0, stream, use_pdl);
  }
  const int stages = variant == 1 ? 4 : (variant == 3 ? 3 : 2);
  const size_t smem = (size_t)LIN_WARPS * stages * 2 * TS_CHUNK * 2 + (size_t)((a.K + 63) / 64) * 64 * 2 +
                      (size_t)LIN_WARPS * stages * 8 + (a.hist ? STAT_BINS * 4 : 0);
  if (smem > 227 * 1024) return launch_stream<MODE>(a, 0, ctas_per_sm, use_pdl, stream);  // huge K: LDG path
  const int per_sm = variant == 1 ? 1 : max(1, min(ctas_per_sm, (int)((227 * 1024) / (smem + 1024))));
  const int grid = grid_override > 0 ? min(grid_override, max_useful) : max(1, min(sms * per_sm, max_useful));
  StreamArgs b = a;
  if (g_l2_prefetch_mb > 0 && b.l2_pf_chunks == 0) {  // bytes of L2 look-ahead per launch, spread evenly over the warps
    const size_t per_step = (size_t)grid * LIN_WARPS * 2 * TS_CHUNK * 2;
    b.l2_pf_chunks = (int)min((size_t)16, ((size_t)g_l2_prefetch_mb << 20) / per_step);
  }
  if (variant == 1) return launch_pdl(stream_bulk_kernel<MODE, 4>, b, grid, smem, stream, use_pdl);
  if (variant == 3) return launch_pdl(stream_bulk_kernel<MODE, 3>, b, grid, smem, stream, use_pdl);
  return launch_pdl(stream_bulk_kernel<MODE, 2>, b, grid, smem, stream, use_pdl);
}

static int g_default_variant = 2;  // bulk-copy ring x2 stages: fastest in-pipeline (profiles/README.md)

}  // namespace mdi

using namespace mdi;

extern "C" {

void mdi_set_linear_variant(int v) { g_default_variant = v; }
void mdi_set_moe_variant(int v) { g_moe_variant = v; }  // expert passes: 0 = register-streamed, else the bulk-copy ring
void mdi_set_l2_prefetch_mb(int mb) { g_l2_prefetch_mb = mb; }
int mdi_get_linear_variant() { return g_default_variant; }

int mdi_linear_decode(const void* W, const void* W2, const void* bias, const void* bias2, const void* x,
                      const void* norm_w, const void* residual, void* y, const int* ctx, long long x_slot_stride,
                      long long res_slot_stride, long long y_slot_stride, int N, int K, float eps, int unit_offset,
                      int act, int out_fp32, const int* wait_flag, int* status, long long wait_max_cycles,
                      int* signal_flag, unsigned int* done_ctr, int ctas_per_sm, int use_pdl, int variant,
                      unsigned int* hist, unsigned long long* amax, unsigned long long* trace, const float* wscale,
                      const float* wscale2, const int* dep_wait_flag, int* dep_signal_flag, unsigned int* dep_ctr,
                      void* hop_row, long long hop_slot_stride, const void* pf_a, const void* pf_b,
                      unsigned long long pf_bytes, int l2_pf_chunks, const void* hop_pre, long long hop_pre_slot_stride,
                      int hop_pre_elems, const void* norm_b, int layer_norm, cudaStream_t stream) {
  if (K % 8 != 0) return -2;
  if (hop_row && (out_fp32 || !signal_flag || !y)) return -2;
  StreamArgs a{};
  a.W = (const bf16*)W; a.W2 = (const bf16*)W2; a.bias = (const bf16*)bias; a.bias2 = (const bf16*)bias2;
  a.x = (const bf16*)x; a.norm_w = (const bf16*)norm_w; a.residual = (const bf16*)residual; a.y = y; a.ctx = ctx;
  a.norm_b = (const bf16*)norm_b; a.layer_norm = layer_norm;
  a.x_slot_stride = x_slot_stride; a.res_slot_stride = res_slot_stride; a.y_slot_stride = y_slot_stride;
  a.N = N; a.K = K; a.eps = eps; a.unit_offset = unit_offset; a.act = act; a.out_fp32 = out_fp32;
  a.wait = HopWait{wait_flag, status, wait_max_cycles};
  a.signal = HopSignal{signal_flag, done_ctr, status};
  a.n_items = W2 ? N : (N + 1) / 2;
  a.hist = hist; a.amax = amax; a.trace = trace; a.wscale = wscale; a.wscale2 = wscale2;
  a.dep_wait = DepWait{dep_wait_flag, status, wait_max_cycles}; a.dep_signal = DepSignal{dep_signal_flag, dep_ctr};
  a.ctx_early = (use_pdl >> 1) & 1; use_pdl &= 1;  // launch flags: bit 0 = PDL, bit 1 = ctx readable before the wait
  a.hop_row = (bf16*)hop_row; a.hop_slot_stride = hop_slot_stride;
  a.hop_pre = (const bf16*)hop_pre; a.hop_pre_slot_stride = hop_pre_slot_stride; a.hop_pre_elems = hop_pre ? hop_pre_elems : 0;
  if (a.hop_pre_elems % 8 != 0) return -2;
  a.pf_a = (const unsigned char*)pf_a; a.pf_b = (const unsigned char*)pf_b; a.pf_bytes = pf_bytes;
  a.l2_pf_chunks = l2_pf_chunks > 0 ? l2_pf_chunks : 0;
  if (variant < 0) variant = g_default_variant;
  if (W2) return launch_stream<MODE_GATED>(a, variant, ctas_per_sm, use_pdl, stream);
  return launch_stream<MODE_PLAIN>(a, variant, ctas_per_sm, use_pdl, stream);
}

int mdi_qkv_decode(const void* W, const void* bias, const void* x, const void* norm_w, const float* cos,
                   const float* sin, void* q_out, void* kv, const int* ctx, long long x_slot_stride, int K,
                   int n_head, int n_groups, int head_size, int rope_n_elem, int max_seq, float eps,
                   int unit_offset, const int* wait_flag, int* status, long long wait_max_cycles, int ctas_per_sm,
                   int use_pdl, int variant, unsigned long long* trace, const float* wscale, const int* dep_wait_flag,
                   int* dep_signal_flag, unsigned int* dep_ctr, const void* norm_b, int layer_norm, cudaStream_t stream) {
  if (K % 8 != 0 || head_size % 2 != 0 || rope_n_elem % 2 != 0 || rope_n_elem > head_size) return -2;
  StreamArgs a{};
  a.W = (const bf16*)W; a.bias = (const bf16*)bias; a.x = (const bf16*)x; a.norm_w = (const bf16*)norm_w;
  a.norm_b = (const bf16*)norm_b; a.layer_norm = layer_norm;
  a.cos = cos; a.sin = sin; a.q_out = (bf16*)q_out; a.kv = (bf16*)kv; a.ctx = ctx; a.x_slot_stride = x_slot_stride;
  a.K = K; a.N = (n_head + 2 * n_groups) * head_size;
  a.n_head = n_head; a.n_groups = n_groups; a.head_size = head_size; a.rope_n_elem = rope_n_elem;
  a.max_seq = max_seq; a.eps = eps; a.unit_offset = unit_offset;
  a.wait = HopWait{wait_flag, status, wait_max_cycles};
  a.signal = HopSignal{nullptr, nullptr, nullptr};
  a.trace = trace; a.wscale = wscale;
  a.dep_wait = DepWait{dep_wait_flag, status, wait_max_cycles}; a.dep_signal = DepSignal{dep_signal_flag, dep_ctr};
  a.ctx_early = (use_pdl >> 1) & 1; use_pdl &= 1;
  a.n_items = (n_head + 2 * n_groups) * (head_size / 2);
  if (variant < 0) variant = g_default_variant;
  return launch_stream<MODE_QKV>(a, variant, ctas_per_sm, use_pdl, stream);
}

// Router of a mixture-of-experts MLP: norm_2 + [E, K] GEMV + top-k + softmax -> sel[top], wts[top] (device memory).
int mdi_moe_router(const void* Wg, const void* x, const void* norm_w, const void* norm_b, int layer_norm, const int* ctx,
                   long long x_slot_stride, int K, int E, int top, float eps, int unit_offset, int* sel, float* wts,
                   const int* wait_flag, int* status, long long wait_max_cycles, int use_pdl, unsigned long long* trace,
                   cudaStream_t stream) {
  if (K % 8 != 0 || E < 1 || E > MOE_MAX_EXPERTS || top < 1 || top > MOE_MAX_TOP || top > E) return -2;
  RouterArgs a{};
  a.Wg = (const bf16*)Wg; a.x = (const bf16*)x; a.norm_w = (const bf16*)norm_w; a.norm_b = (const bf16*)norm_b;
  a.layer_norm = layer_norm; a.ctx = ctx; a.x_slot_stride = x_slot_stride; a.K = K; a.E = E; a.top = top; a.eps = eps;
  a.unit_offset = unit_offset; a.sel = sel; a.wts = wts; a.wait = HopWait{wait_flag, status, wait_max_cycles}; a.trace = trace;
  const size_t smem = (size_t)((K + 63) / 64) * 64 * sizeof(bf16);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(moe_router_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(1);
  cfg.blockDim = dim3(LIN_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  return (int)cudaLaunchKernelEx(&cfg, moe_router_kernel, a);
}

// One expert pass of a routed token: gate/up (w2_ptrs != null) or down (+ routing weight, running sum, residual, hop).
int mdi_moe_linear_decode(const void* w_ptrs, const void* w2_ptrs, const int* sel, const float* wts, int k, const void* prev,
                          const void* x, const void* norm_w, const void* norm_b, int layer_norm, const void* residual, void* y,
                          const int* ctx, long long x_slot_stride, long long res_slot_stride, long long y_slot_stride, int N,
                          int K, float eps, int unit_offset, int act, int* status, int* signal_flag, unsigned int* done_ctr,
                          void* hop_row, long long hop_slot_stride, int ctas_per_sm, int use_pdl, unsigned long long* trace,
                          int sel_early, cudaStream_t stream) {
  if (K % 8 != 0 || !w_ptrs || !sel || !wts || k < 0 || k >= MOE_MAX_TOP) return -2;
  if (hop_row && (!signal_flag || !y)) return -2;
  StreamArgs a{};
  a.x = (const bf16*)x; a.norm_w = (const bf16*)norm_w; a.norm_b = (const bf16*)norm_b; a.layer_norm = layer_norm;
  a.residual = (const bf16*)residual; a.y = y; a.ctx = ctx;
  a.x_slot_stride = x_slot_stride; a.res_slot_stride = res_slot_stride; a.y_slot_stride = y_slot_stride;
  a.N = N; a.K = K; a.eps = eps; a.unit_offset = unit_offset; a.act = act;
  a.signal = HopSignal{signal_flag, done_ctr, status};
  a.n_items = w2_ptrs ? N : (N + 1) / 2;
  a.trace = trace; a.hop_row = (bf16*)hop_row; a.hop_slot_stride = hop_slot_stride;
  MoeArgs m{};
  m.w = (const bf16* const*)w_ptrs; m.w2 = (const bf16* const*)w2_ptrs; m.sel = sel; m.wts = wts; m.k = k;
  m.prev = (const bf16*)prev; m.sel_early = sel_early;
  if (ctas_per_sm <= 0) ctas_per_sm = 3;
  const int max_useful = (a.n_items + LIN_WARPS - 1) / LIN_WARPS;
  int grid = max(1, min(num_sms() * min(ctas_per_sm, 3), max_useful));
  size_t smem = (size_t)((K + 63) / 64) * 64 * sizeof(bf16);
  const size_t smem_ring = (size_t)LIN_WARPS * 2 * 2 * TS_CHUNK * 2 + smem + (size_t)LIN_WARPS * 2 * 8;
  const bool ring = g_moe_variant != 0 && smem_ring <= 227 * 1024;
  if (ring) {
    smem = smem_ring;
    grid = max(1, min(num_sms() * max(1, min(min(ctas_per_sm, 3), (int)((227 * 1024) / (smem + 1024)))), max_useful));
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(LIN_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  if (ring) {
    auto kern = w2_ptrs ? moe_bulk_kernel<MODE_GATED, 2> : moe_bulk_kernel<MODE_PLAIN, 2>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    return (int)cudaLaunchKernelEx(&cfg, kern, a, m);
  }
  if (w2_ptrs) {
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(moe_stream_kernel<MODE_GATED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return (int)e;
    }
    return (int)cudaLaunchKernelEx(&cfg, moe_stream_kernel<MODE_GATED>, a, m);
  }
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(moe_stream_kernel<MODE_PLAIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
  }
  return (int)cudaLaunchKernelEx(&cfg, moe_stream_kernel<MODE_PLAIN>, a, m);
}

}  // extern "C"

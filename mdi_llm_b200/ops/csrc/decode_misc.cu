// Small decode-path kernels: embedding gather, stand-alone RMSNorm, sampling, step bookkeeping.
//
//  * embed        — SURVEY K1 (submodels.py:210-212): wte row gather (+ sqrt(C) scale for Gemma,
//                   + learned position row for the GPT-2 family); the token id is read on device
//                   (written by the sampler of the previous round) — no host round trip.
//  * rmsnorm_rows — K2 for multi-row inputs (prefill); decode fuses the norm into its consumer.
//  * sample       — K13 (model.py:67-90): greedy arg-max, or top-k -> temperature -> softmax ->
//                   one multinomial draw, entirely on device with a counter-based RNG; replaces
//                   topk + scatter + softmax + multinomial (≈6 launches + a D2H sync per token,
//                   gptserver.py:933-949).  Exact k-th-largest via 4-pass radix select.
//  * advance_step — device-driven pipeline: derive {slot, pos, wait, signal} for the next graph
//                   replay from a device-resident counter (round-robin over the samples).
#include "common.cuh"

namespace mdi {

__global__ void embed_kernel(const bf16* __restrict__ wte, const bf16* __restrict__ wpe, const int* __restrict__ tokens,
                             long long tok_slot_stride, const int* __restrict__ ctx, bf16* __restrict__ x,
                             long long x_slot_stride, int C, float scale) {
  pdl_wait_prior();
  pdl_launch_dependents();
  const int slot = ctx[MDI_CTX_SLOT], pos = ctx[MDI_CTX_POS];
  const int tok = tokens ? tokens[(size_t)slot * tok_slot_stride + pos] : ctx[MDI_CTX_TOKEN];
  const bf16* row = wte + (size_t)tok * C;
  bf16* dst = x + (size_t)slot * x_slot_stride;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < C; i += gridDim.x * blockDim.x) {
    float v = __bfloat162float(row[i]);
    if (scale != 1.f) v = round_bf16(v * scale);
    if (wpe) v = v + __bfloat162float(wpe[(size_t)pos * C + i]);
    dst[i] = __float2bfloat16_rn(v);
  }
}

// y[r] = bf16(x[r] * rstd) * w   — one CTA per row
__global__ void __launch_bounds__(256) rmsnorm_rows_kernel(const bf16* __restrict__ x, const bf16* __restrict__ w,
                                                           bf16* __restrict__ y, int C, float eps, int unit_offset) {
  __shared__ float red[8];
  const bf16* xr = x + (size_t)blockIdx.x * C;
  bf16* yr = y + (size_t)blockIdx.x * C;
  float ss = 0.f;
  for (int i = threadIdx.x; i < C; i += 256) { float f = __bfloat162float(xr[i]); ss = fmaf(f, f, ss); }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float rstd = rsqrtf(tot / (float)C + eps);
  for (int i = threadIdx.x; i < C; i += 256) {
    float xn = round_bf16(__bfloat162float(xr[i]) * rstd);
    float ww = __bfloat162float(w[i]);
    if (unit_offset) ww = round_bf16(1.f + ww);
    yr[i] = __float2bfloat16_rn(xn * ww);
  }
}

// ------------------------------------------------------------------------------------------------
// sampling
constexpr int SMP_THREADS = 1024;
constexpr int SMP_KMAX = 1024;

__device__ __forceinline__ uint32_t float_key(float f) {  // monotone: larger float -> larger key
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}


// ---- general sampler (any top-k, nucleus top-p) -------------------------------------------------------------
// One CTA of SMP_THREADS threads over the whole vocabulary.  Semantics of the eager `sample()`
// (models/gpt.py, reference model.py:42-90): crop to the k largest logits, divide by the temperature, drop the
// low-probability tail whose ascending cumulative probability is <= 1 - top_p (i.e. keep a token iff the
// probability mass strictly above it is < top_p), renormalise, one multinomial draw.  Both crops are radix
// selections over the orderable 32-bit logit key: by COUNT for top-k, by probability MASS for top-p; ties of the
// threshold key are kept as a group.  Histogram updates are warp-aggregated (most logits share their top bits).
__device__ __forceinline__ void hist_add_u(unsigned int* hist, bool on, unsigned int digit) {
  const unsigned int peers = __match_any_sync(0xffffffffu, on ? digit : 0xffffffffu);
  if (on && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(hist + digit, (unsigned)__popc(peers));
}
__device__ __forceinline__ void hist_add_f(float* hist, bool on, unsigned int digit, float w) {
  const unsigned int peers = __match_any_sync(0xffffffffu, on ? digit : 0xffffffffu);
  float tot = 0.f;
  for (unsigned int m = peers; m; m &= m - 1) tot += __shfl_sync(peers, w, __ffs(m) - 1);
  if (on && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(hist + digit, tot);
}

// scratch: 256 u32 + 256 f32 + 40 f32 in shared memory
__device__ int sample_full(const float* __restrict__ lg, int V, int k, float top_p, float inv_t, float u01,
                           unsigned int* hist, float* fhist, float* red) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __shared__ unsigned int sf_prefix, sf_kleft;
  __shared__ float sf_above, sf_val;
  __shared__ int sf_pick;
  // ---- max logit ------------------------------------------------------------------------------------------
  float mx = -INFINITY;
  for (int i = tid; i < V; i += SMP_THREADS) mx = fmaxf(mx, __ldcg(lg + i));
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (warp == 0) { const float t = warp_max(red[lane]); if (lane == 0) sf_val = t; }
  __syncthreads();
  mx = sf_val;
  __syncthreads();
  // ---- top-k: key of the k-th largest (by count) ------------------------------------------------------------
  uint32_t thr = 0u;
  if (k > 0 && k < V) {
    if (tid == 0) { sf_prefix = 0; sf_kleft = (unsigned)k; }
    __syncthreads();
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      if (tid < 256) hist[tid] = 0;
      __syncthreads();
      const unsigned int prefix = sf_prefix;
      const unsigned int mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
      for (int i0 = 0; i0 < V; i0 += SMP_THREADS) {
        const int i = i0 + tid;
        const uint32_t key = i < V ? float_key(__ldcg(lg + i)) : 0u;
        hist_add_u(hist, i < V && (key & mask) == prefix, (key >> shift) & 0xffu);
      }
      __syncthreads();
      if (tid == 0) {
        unsigned int left = sf_kleft, cum = 0;
        int b = 255;
        for (; b > 0; --b) {
          if (cum + hist[b] >= left) break;
          cum += hist[b];
        }
        sf_kleft = left - cum;
        sf_prefix = prefix | ((unsigned)b << shift);
      }
      __syncthreads();
    }
    thr = sf_prefix;
    __syncthreads();
  }
  // ---- probability mass of the kept set ------------------------------------------------------------------------
  auto mass_ge = [&](uint32_t t) -> float {
    float s = 0.f;
    for (int i = tid; i < V; i += SMP_THREADS) {
      const float v = __ldcg(lg + i);
      if (float_key(v) >= t) s += __expf((v - mx) * inv_t);
    }
    s = warp_sum(s);
    __syncthreads();
    if (lane == 0) red[warp] = s;
    __syncthreads();
    if (warp == 0) { const float t2 = warp_sum(red[lane]); if (lane == 0) sf_val = t2; }
    __syncthreads();
    return sf_val;
  };
  float total = mass_ge(thr);
  // ---- top-p: smallest key K with mass(keys > K) < top_p * total ------------------------------------------------
  if (top_p < 1.f) {
    const float P = top_p * total;
    if (tid == 0) { sf_prefix = 0; sf_above = 0.f; }
    __syncthreads();
    for (int pass = 0; pass < 4; ++pass) {
      const int shift = 24 - 8 * pass;
      if (tid < 256) fhist[tid] = 0.f;
      __syncthreads();
      const unsigned int prefix = sf_prefix;
      const unsigned int mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
      for (int i0 = 0; i0 < V; i0 += SMP_THREADS) {
        const int i = i0 + tid;
        const float v = i < V ? __ldcg(lg + i) : 0.f;
        const uint32_t key = float_key(v);
        const bool on = i < V && key >= thr && (key & mask) == prefix;
        hist_add_f(fhist, on, (key >> shift) & 0xffu, on ? __expf((v - mx) * inv_t) : 0.f);
      }
      __syncthreads();
      if (tid == 0) {
        float above = sf_above;  // mass of every key above the current prefix range
        int b = 255;
        // descend while the mass strictly above bin b's successor still leaves room: stop at the lowest bin whose
        // "mass of higher bins" is < P
        for (; b > 0; --b) {
          if (above + fhist[b] >= P) break;  // bins below b would have mass-above >= P: dropped
          above += fhist[b];
        }
        sf_above = above;
        sf_prefix = prefix | ((unsigned)b << shift);
      }
      __syncthreads();
    }
    uint32_t thr_p = sf_prefix;
    __syncthreads();
    if (thr_p > thr) thr = thr_p;
    total = mass_ge(thr);
  }
  // ---- one draw in vocabulary order (a fixed seed gives a fixed token) ---------------------------------------------
  const int chunk = (V + SMP_THREADS - 1) / SMP_THREADS;
  const int lo = tid * chunk, hi = min(V, lo + chunk);
  float mine = 0.f;
  for (int i = lo; i < hi; ++i) {
    const float v = __ldcg(lg + i);
    if (float_key(v) >= thr) mine += __expf((v - mx) * inv_t);
  }
  float sc = mine;  // inclusive scan over threads
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_up_sync(0xffffffffu, sc, o); if (lane >= o) sc += t; }
  __syncthreads();
  if (lane == 31) red[warp] = sc;
  __syncthreads();
  if (warp == 0) {
    float w = red[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
    red[lane] = w;
  }
  __syncthreads();
  const float incl = sc + (warp > 0 ? red[warp - 1] : 0.f), excl = incl - mine;
  const float target = u01 * red[31];
  if (tid == 0) sf_pick = -1;
  __syncthreads();
  if (mine > 0.f && target >= excl && target < incl) {
    float run = excl;
    int pick = -1;
    for (int i = lo; i < hi; ++i) {
      const float v = __ldcg(lg + i);
      if (float_key(v) >= thr) {
        pick = i;  // last kept element of the chunk absorbs round-off
        run += __expf((v - mx) * inv_t);
        if (target < run) break;
      }
    }
    sf_pick = pick;
  }
  __syncthreads();
  if (sf_pick < 0) {  // round-off at the upper end (or an all -inf row): fall back to the arg-max
    float best = -INFINITY; int bi = 0x7fffffff;
    for (int i = tid; i < V; i += SMP_THREADS) { const float v = __ldcg(lg + i); if (v > best || (v == best && i < bi)) { best = v; bi = i; } }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, best, o); const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    __syncthreads();
    if (lane == 0) { red[warp] = best; hist[warp] = (unsigned)bi; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 0; w < SMP_THREADS / 32; ++w) if (red[w] > best || (red[w] == best && (int)hist[w] < bi)) { best = red[w]; bi = (int)hist[w]; }
      sf_pick = bi;
    }
    __syncthreads();
  }
  return sf_pick;
}

struct SampleArgs {
  const float* logits;  // [V] (+ slot * logits_slot_stride)
  long long logits_slot_stride;
  int* tokens;          // sampled id stored at tokens[slot * tok_slot_stride + pos]
  long long tok_slot_stride;
  int* last_token;      // optional: [n_slots] latest token per slot (host-visible mirror)
  const int* ctx;
  int V;                // sample among the first V logits (true vocab may be < padded)
  int top_k;            // <= 0: no top-k crop
  float temperature;    // <= 0 with greedy != 0: arg-max
  int greedy;
  unsigned long long seed;
  float top_p;          // >= 1: no nucleus crop
};

__global__ void __launch_bounds__(SMP_THREADS) sample_kernel(const SampleArgs a) {
  __shared__ unsigned int hist[256];
  __shared__ float cand_val[SMP_KMAX];
  __shared__ int cand_idx[SMP_KMAX];
  __shared__ float sorted_val[SMP_KMAX];
  __shared__ int sorted_idx[SMP_KMAX];
  __shared__ float red_f[32];
  __shared__ int red_i[32];
  __shared__ unsigned int sh_prefix, sh_kleft, sh_count, sh_ngreater;
  __shared__ float sh_max, sh_sum;
  __shared__ int sh_pick;

  pdl_wait_prior();
  pdl_launch_dependents();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slot = a.ctx[MDI_CTX_SLOT], pos = a.ctx[MDI_CTX_POS];
  const float* lg = a.logits + (size_t)slot * a.logits_slot_stride;
  const int V = a.V;

  if (a.greedy) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int i = tid; i < V; i += SMP_THREADS) {
      float v = lg[i];
      if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if (lane == 0) { red_f[warp] = best; red_i[warp] = bi; }
    __syncthreads();
    if (warp == 0) {
      best = red_f[lane]; bi = red_i[lane];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        float ov = __shfl_xor_sync(0xffffffffu, best, o);
        int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
      }
      if (lane == 0) {
        a.tokens[(size_t)slot * a.tok_slot_stride + pos] = bi;
        if (a.last_token) a.last_token[slot] = bi;
      }
    }
    return;
  }

  int k = (a.top_k > 0 && a.top_k < V) ? a.top_k : V;
  if (k > SMP_KMAX || a.top_p < 1.f) {  // beyond the candidate-list path: whole-vocabulary sampler
    __shared__ float fh[256];
    const uint64_t r0 = splitmix64(a.seed ^ splitmix64(((uint64_t)(uint32_t)slot << 32) | (uint32_t)pos));
    const int tok = sample_full(lg, V, k < V ? k : 0, a.top_p, a.temperature > 0.f ? 1.f / a.temperature : 1.f,
                                (float)((r0 >> 40) * (1.0 / 16777216.0)), hist, fh, red_f);
    if (tid == 0) {
      a.tokens[(size_t)slot * a.tok_slot_stride + pos] = tok;
      if (a.last_token) a.last_token[slot] = tok;
    }
    return;
  }
  // ---- radix select: key of the k-th largest logit ---------------------------------------------
  if (tid == 0) { sh_prefix = 0; sh_kleft = (unsigned)k; sh_ngreater = 0; }
  __syncthreads();
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    const unsigned int prefix = sh_prefix;
    const unsigned int mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    for (int i = tid; i < V; i += SMP_THREADS) {
      const uint32_t key = float_key(lg[i]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned int left = sh_kleft, cum = 0;
      int b = 255;
      for (; b > 0; --b) {
        if (cum + hist[b] >= left) break;
        cum += hist[b];
      }
      sh_ngreater += cum;
      sh_kleft = left - cum;
      sh_prefix = prefix | ((unsigned)b << shift);
    }
    __syncthreads();
  }
  const uint32_t thr = sh_prefix;        // exact key of the k-th largest
  const unsigned int n_ties = sh_kleft;  // how many elements equal to thr belong to the top-k
  if (tid == 0) sh_count = 0;
  __syncthreads();
  // ---- collect the top-k: first everything strictly greater, then the needed ties -------------
  for (int i = tid; i < V; i += SMP_THREADS) {
    const float v = lg[i];
    if (float_key(v) > thr) {
      unsigned int s = atomicAdd(&sh_count, 1u);
      if (s < SMP_KMAX) { cand_val[s] = v; cand_idx[s] = i; }
    }
  }
  __syncthreads();
  const unsigned int n_greater = min(sh_count, (unsigned)SMP_KMAX);
  __syncthreads();
  if (tid == 0) sh_count = 0;
  __syncthreads();
  for (int i = tid; i < V; i += SMP_THREADS) {
    const float v = lg[i];
    if (float_key(v) == thr) {
      unsigned int s = atomicAdd(&sh_count, 1u);
      if (s < n_ties && n_greater + s < SMP_KMAX) { cand_val[n_greater + s] = v; cand_idx[n_greater + s] = i; }
    }
  }
  __syncthreads();
  const int n = (int)min(n_greater + n_ties, (unsigned)SMP_KMAX);
  // ---- order candidates by vocabulary index so a fixed seed gives a fixed token ---------------
  if (tid < n) {
    const int my = cand_idx[tid];
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += (cand_idx[j] < my);
    sorted_idx[rank] = my;
    sorted_val[rank] = cand_val[tid];
  }
  __syncthreads();
  // ---- softmax(logit / T) over the candidates ------------------------------------------------------
  const float inv_t = a.temperature > 0.f ? 1.f / a.temperature : 1.f;
  float v = tid < n ? sorted_val[tid] * inv_t : -INFINITY;
  float mx = warp_max(v);
  if (lane == 0) red_f[warp] = mx;
  __syncthreads();
  if (warp == 0) { float t = warp_max(red_f[lane]); if (lane == 0) sh_max = t; }
  __syncthreads();
  float e = tid < n ? __expf(v - sh_max) : 0.f;
  // inclusive scan of e over the block (warp scans + scan of warp totals)
  float sc = e;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { float t = __shfl_up_sync(0xffffffffu, sc, o); if (lane >= o) sc += t; }
  __syncthreads();
  if (lane == 31) red_f[warp] = sc;
  __syncthreads();
  if (warp == 0) {
    float w = red_f[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { float t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
    red_f[lane] = w;
    if (lane == 31) sh_sum = w;
  }
  __syncthreads();
  const float incl = sc + (warp > 0 ? red_f[warp - 1] : 0.f);
  const float excl = incl - e;
  // one uniform draw from a counter-based generator keyed by (seed, slot, pos)
  const uint64_t r = splitmix64(a.seed ^ splitmix64(((uint64_t)(uint32_t)slot << 32) | (uint32_t)pos));
  const float u = (float)((r >> 40) * (1.0 / 16777216.0)) * sh_sum;
  if (tid == 0) sh_pick = sorted_idx[n - 1];  // guard against round-off at the upper end
  __syncthreads();
  if (tid < n && e > 0.f && u >= excl && u < incl) sh_pick = sorted_idx[tid];
  __syncthreads();
  if (tid == 0) {
    a.tokens[(size_t)slot * a.tok_slot_stride + pos] = sh_pick;
    if (a.last_token) a.last_token[slot] = sh_pick;
  }
}

// ------------------------------------------------------------------------------------------------
// Fast sampler (the decode path): the lm_head epilogue has already histogrammed the top 12 key
// bits of every logit (and tracked the arg-max), so sampling is two short kernels:
//   filter (many CTAs): find the histogram bin holding the k-th largest logit, append every
//                       logit at or above that bin to a candidate list (a few hundred entries);
//   final  (one CTA)  : exact top-k among the candidates (radix select in shared memory),
//                       temperature softmax, one draw; resets the statistics for the next step.
constexpr int SF_BINS = 4096;
constexpr int SF_CAND_MAX = 8192;

struct SampleFastArgs {
  const float* logits;
  unsigned int* hist;           // [4096]
  unsigned long long* amax;     // [1]
  unsigned int* cand_count;     // [1]
  float* cand_val;              // [SF_CAND_MAX]
  int* cand_idx;                // [SF_CAND_MAX]
  int* tokens; long long tok_slot_stride; int* last_token;
  const int* ctx;
  int V, top_k; float temperature; int greedy; unsigned long long seed;
  float top_p;                  // >= 1: no nucleus crop
  unsigned long long* tok_ts;   // optional [n_slots, ts_slot_stride]: %globaltimer when the token was produced
  long long ts_slot_stride;
  int* status;                  // status words of the stage (bit 4 of [0]: candidate overflow, handled exactly)
};

__global__ void __launch_bounds__(256) sample_filter_kernel(const SampleFastArgs a) {
  __shared__ unsigned int tsum[256];
  __shared__ int sh_bin;
  pdl_wait_prior();
  pdl_launch_dependents();
  if (a.greedy) return;
  const int tid = threadIdx.x;
  const int k = (a.top_k > 0 && a.top_k < a.V) ? min(a.top_k, 1024) : min(a.V, 1024);
  // threshold bin: highest b with sum_{j >= b} hist[j] >= k   (thread t owns bins [16t, 16t+16))
  unsigned int loc[16], s = 0;
#pragma unroll
  for (int j = 0; j < 16; ++j) { loc[j] = __ldcg(a.hist + tid * 16 + j); s += loc[j]; }
  tsum[tid] = s;
  if (tid == 0) sh_bin = 0;
  __syncthreads();
  unsigned int above = 0;  // population of all bins owned by higher threads
  for (int t = tid + 1; t < 256; ++t) above += tsum[t];
  if (above < (unsigned)k && above + s >= (unsigned)k) {
    unsigned int cum = above;
    int j = 15;
    for (; j > 0; --j) {
      if (cum + loc[j] >= (unsigned)k) break;
      cum += loc[j];
    }
    sh_bin = tid * 16 + j;
  }
  __syncthreads();
  const uint32_t bin = (uint32_t)sh_bin;
  for (int i = blockIdx.x * blockDim.x + tid; i < a.V; i += gridDim.x * blockDim.x) {
    const float v = __ldcg(a.logits + i);
    if ((float_key(v) >> 20) >= bin) {
      const unsigned int slot = atomicAdd(a.cand_count, 1u);
      if (slot < SF_CAND_MAX) { a.cand_val[slot] = v; a.cand_idx[slot] = i; }
    }
  }
}

__global__ void __launch_bounds__(SMP_THREADS) sample_final_kernel(const SampleFastArgs a) {
  extern __shared__ __align__(16) unsigned char sf_smem[];
  float* cv = reinterpret_cast<float*>(sf_smem);             // [SF_CAND_MAX]
  int* ci = reinterpret_cast<int*>(sf_smem) + SF_CAND_MAX;  // [SF_CAND_MAX]
  __shared__ unsigned int hist[256];
  __shared__ float top_val[SMP_KMAX], sorted_val[SMP_KMAX];
  __shared__ int top_idx[SMP_KMAX], sorted_idx[SMP_KMAX];
  __shared__ float red_f[32];
  __shared__ unsigned int sh_prefix, sh_kleft, sh_count;
  __shared__ float sh_max, sh_sum;
  __shared__ int sh_pick;
  pdl_wait_prior();
  pdl_launch_dependents();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int slot = a.ctx[MDI_CTX_SLOT], pos = a.ctx[MDI_CTX_POS];

  if (a.greedy) {
    if (tid == 0) {
      const unsigned long long best = *a.amax;
      const int tok = (int)(0xffffffffu - (uint32_t)(best & 0xffffffffull));
      a.tokens[(size_t)slot * a.tok_slot_stride + pos] = tok;
      if (a.last_token) a.last_token[slot] = tok;
      if (a.tok_ts) a.tok_ts[(size_t)slot * a.ts_slot_stride + pos] = globaltimer_ns();
      *a.amax = 0ull;
    }
    for (int i = tid; i < SF_BINS; i += SMP_THREADS) a.hist[i] = 0;
    return;
  }
  const unsigned int n_cand_raw = __ldcg(a.cand_count);
  {
    const int kk = (a.top_k > 0 && a.top_k < a.V) ? a.top_k : a.V;
    if (n_cand_raw > (unsigned)SF_CAND_MAX || kk > SMP_KMAX || a.top_p < 1.f) {
      // nucleus sampling, k beyond the candidate list, or a logit distribution so flat that the threshold bin
      // overflowed the list: exact whole-vocabulary sampler (slower, never wrong)
      if (tid == 0 && n_cand_raw > (unsigned)SF_CAND_MAX && a.status) atomicOr(a.status, 4);
      const uint64_t r0 = splitmix64(a.seed ^ splitmix64(((uint64_t)(uint32_t)slot << 32) | (uint32_t)pos));
      const int tok = sample_full(a.logits, a.V, kk < a.V ? kk : 0, a.top_p, a.temperature > 0.f ? 1.f / a.temperature : 1.f,
                                  (float)((r0 >> 40) * (1.0 / 16777216.0)), hist, reinterpret_cast<float*>(top_val), red_f);
      if (tid == 0) {
        a.tokens[(size_t)slot * a.tok_slot_stride + pos] = tok;
        if (a.last_token) a.last_token[slot] = tok;
        if (a.tok_ts) a.tok_ts[(size_t)slot * a.ts_slot_stride + pos] = globaltimer_ns();
        *a.cand_count = 0;
        *a.amax = 0ull;
      }
      for (int i = tid; i < SF_BINS; i += SMP_THREADS) a.hist[i] = 0;
      return;
    }
  }
  const int n_cand = (int)min(n_cand_raw, (unsigned)SF_CAND_MAX);
  for (int i = tid; i < n_cand; i += SMP_THREADS) { cv[i] = __ldcg(a.cand_val + i); ci[i] = __ldcg(a.cand_idx + i); }
  int k = (a.top_k > 0 && a.top_k < a.V) ? a.top_k : a.V;
  k = min(min(k, SMP_KMAX), n_cand);
  if (tid == 0) { sh_prefix = 0; sh_kleft = (unsigned)k; }
  __syncthreads();
  for (int pass = 0; pass < 4; ++pass) {  // exact k-th largest among the candidates
    const int shift = 24 - 8 * pass;
    if (tid < 256) hist[tid] = 0;
    __syncthreads();
    const unsigned int prefix = sh_prefix;
    const unsigned int mask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    for (int i = tid; i < n_cand; i += SMP_THREADS) {
      const uint32_t key = float_key(cv[i]);
      if ((key & mask) == prefix) atomicAdd(&hist[(key >> shift) & 0xffu], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      unsigned int left = sh_kleft, cum = 0;
      int b = 255;
      for (; b > 0; --b) {
        if (cum + hist[b] >= left) break;
        cum += hist[b];
      }
      sh_kleft = left - cum;
      sh_prefix = prefix | ((unsigned)b << shift);
    }
    __syncthreads();
  }
  const uint32_t thr = sh_prefix;
  const unsigned int n_ties = sh_kleft;
  if (tid == 0) sh_count = 0;
  __syncthreads();
  for (int i = tid; i < n_cand; i += SMP_THREADS)
    if (float_key(cv[i]) > thr) {
      const unsigned int s = atomicAdd(&sh_count, 1u);
      if (s < SMP_KMAX) { top_val[s] = cv[i]; top_idx[s] = ci[i]; }
    }
  __syncthreads();
  const unsigned int n_greater = min(sh_count, (unsigned)SMP_KMAX);
  __syncthreads();
  if (tid == 0) sh_count = 0;
  __syncthreads();
  for (int i = tid; i < n_cand; i += SMP_THREADS)
    if (float_key(cv[i]) == thr) {
      const unsigned int s = atomicAdd(&sh_count, 1u);
      if (s < n_ties && n_greater + s < SMP_KMAX) { top_val[n_greater + s] = cv[i]; top_idx[n_greater + s] = ci[i]; }
    }
  __syncthreads();
  const int n = (int)min(n_greater + n_ties, (unsigned)SMP_KMAX);
  if (tid < n) {  // order by vocabulary index: a fixed seed gives a fixed token
    const int my = top_idx[tid];
    int rank = 0;
    for (int j = 0; j < n; ++j) rank += (top_idx[j] < my);
    sorted_idx[rank] = my;
    sorted_val[rank] = top_val[tid];
  }
  __syncthreads();
  const float inv_t = a.temperature > 0.f ? 1.f / a.temperature : 1.f;
  const float v = tid < n ? sorted_val[tid] * inv_t : -INFINITY;
  const float mx = warp_max(v);
  if (lane == 0) red_f[warp] = mx;
  __syncthreads();
  if (warp == 0) { const float t = warp_max(red_f[lane]); if (lane == 0) sh_max = t; }
  __syncthreads();
  const float e = tid < n ? __expf(v - sh_max) : 0.f;
  float sc = e;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_up_sync(0xffffffffu, sc, o); if (lane >= o) sc += t; }
  __syncthreads();
  if (lane == 31) red_f[warp] = sc;
  __syncthreads();
  if (warp == 0) {
    float w = red_f[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const float t = __shfl_up_sync(0xffffffffu, w, o); if (lane >= o) w += t; }
    red_f[lane] = w;
    if (lane == 31) sh_sum = w;
  }
  __syncthreads();
  const float incl = sc + (warp > 0 ? red_f[warp - 1] : 0.f);
  const float excl = incl - e;
  const uint64_t r = splitmix64(a.seed ^ splitmix64(((uint64_t)(uint32_t)slot << 32) | (uint32_t)pos));
  const float u = (float)((r >> 40) * (1.0 / 16777216.0)) * sh_sum;
  if (tid == 0) sh_pick = n > 0 ? sorted_idx[n - 1] : 0;
  __syncthreads();
  if (tid < n && e > 0.f && u >= excl && u < incl) sh_pick = sorted_idx[tid];
  __syncthreads();
  if (tid == 0) {
    a.tokens[(size_t)slot * a.tok_slot_stride + pos] = sh_pick;
    if (a.last_token) a.last_token[slot] = sh_pick;
    if (a.tok_ts) a.tok_ts[(size_t)slot * a.ts_slot_stride + pos] = globaltimer_ns();
    *a.cand_count = 0;
    *a.amax = 0ull;
  }
  for (int i = tid; i < SF_BINS; i += SMP_THREADS) a.hist[i] = 0;  // statistics ready for the next token
}

// ------------------------------------------------------------------------------------------------
// Device-driven schedule.  state[0] = global step counter; per slot: pos[slot].
// Step t serves slot = t % n_slots at round = t / n_slots.
//   secondary : wait = round + 1 (message `round` has arrived), signal = round + 1
//   starter   : head waits `round` (the hidden state that came back), forward signals round + 1
// (round 0 is the prefill, handled outside the graph; decode rounds start at first_round.)
__global__ void advance_step_kernel(int* __restrict__ ctx, int* __restrict__ state, int* __restrict__ pos_arr,
                                    int n_slots, int is_starter) {
  // Release the next kernel at once: it becomes resident and prefetches its weights while the previous step's last
  // kernel is still draining.  Safe: its griddepcontrol.wait returns only after THIS grid has completed, and this
  // grid writes ctx only after the previous step has completed (the wait below) — the chain stays transitive.
  pdl_launch_dependents();
  pdl_wait_prior();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const int t = state[0];
    const int slot = t % n_slots;
    const int round = state[1] + t / n_slots;  // state[1] = first decode round (normally 1)
    ctx[MDI_CTX_SLOT] = slot;
    ctx[MDI_CTX_POS] = pos_arr[slot];
    ctx[MDI_CTX_WAIT] = is_starter ? round : round + 1;
    ctx[MDI_CTX_SIGNAL] = round + 1;
    ctx[MDI_CTX_STEP] = t;
    pos_arr[slot] += 1;
    state[0] = t + 1;
  }
}

__global__ void wait_flag_kernel(const int* flag, const int* ctx, int* status, long long max_cycles) {
  if (threadIdx.x == 0 && blockIdx.x == 0) hop_wait_one(flag + ctx[MDI_CTX_SLOT], ctx[MDI_CTX_WAIT], status, max_cycles);
}
__global__ void set_flag_kernel(int* flag, const int* ctx) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    __threadfence_system();
    st_release_sys(flag + ctx[MDI_CTX_SLOT], ctx[MDI_CTX_SIGNAL]);
  }
}
// %globaltimer -> *dst (time base of the per-token device timeline)
__global__ void stamp_kernel(unsigned long long* dst) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *dst = globaltimer_ns();
}
// copy n 16-byte vectors (prefill hop payload: T x C hidden states) to a (peer) destination
__global__ void copy_vec_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t nvec) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
}

// Prefill hop: copy + publish in ONE kernel.  Every writer fences at system scope before the
// last CTA releases the flag — a separate "set flag" kernel after a copy kernel is NOT enough:
// nothing orders another kernel's peer stores (spread over many NVLink lanes) before the flag store.
__global__ void copy_signal_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t nvec,
                                   HopSignal sig, const int* __restrict__ ctx) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
  hop_signal(sig, ctx);
}

}  // namespace mdi

using namespace mdi;

static int launch_small(const void* kern, dim3 grid, dim3 block, void** args, int use_pdl, cudaStream_t stream) {
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  return (int)cudaLaunchKernelExC(&cfg, kern, args);
}

extern "C" {

int mdi_embed(const void* wte, const void* wpe, const int* tokens, long long tok_slot_stride, const int* ctx, void* x,
              long long x_slot_stride, int C, float scale, int use_pdl, cudaStream_t stream) {
  void* args[] = {&wte, &wpe, &tokens, &tok_slot_stride, &ctx, &x, &x_slot_stride, &C, &scale};
  int threads = 256, blocks = (C + threads - 1) / threads;
  return launch_small((const void*)embed_kernel, dim3(blocks), dim3(threads), args, use_pdl, stream);
}

int mdi_rmsnorm_rows(const void* x, const void* w, void* y, int rows, int C, float eps, int unit_offset,
                     cudaStream_t stream) {
  rmsnorm_rows_kernel<<<rows, 256, 0, stream>>>((const bf16*)x, (const bf16*)w, (bf16*)y, C, eps, unit_offset);
  return (int)cudaGetLastError();
}

int mdi_sample(const float* logits, long long logits_slot_stride, int* tokens, long long tok_slot_stride,
               int* last_token, const int* ctx, int V, int top_k, float temperature, int greedy,
               unsigned long long seed, int use_pdl, float top_p, cudaStream_t stream) {
  SampleArgs a{logits, logits_slot_stride, tokens, tok_slot_stride, last_token, ctx, V, top_k, temperature, greedy, seed, top_p};
  void* args[] = {&a};
  return launch_small((const void*)sample_kernel, dim3(1), dim3(SMP_THREADS), args, use_pdl, stream);
}

// scratch: u32 hist[4096] | u64 amax | u32 cand_count | pad | f32 cand_val[8192] | i32 cand_idx[8192]
size_t mdi_sample_scratch_bytes() { return 4096 * 4 + 8 + 8 + 8192 * 4 + 8192 * 4; }

int mdi_sample_fast(const float* logits, void* scratch, int* tokens, long long tok_slot_stride, int* last_token,
                    const int* ctx, int V, int top_k, float temperature, int greedy, unsigned long long seed,
                    int use_pdl, float top_p, unsigned long long* tok_ts, long long ts_slot_stride, int* status,
                    cudaStream_t stream) {
  char* base = (char*)scratch;
  SampleFastArgs a;
  a.logits = logits;
  a.hist = (unsigned int*)base;
  a.amax = (unsigned long long*)(base + 4096 * 4);
  a.cand_count = (unsigned int*)(base + 4096 * 4 + 8);
  a.cand_val = (float*)(base + 4096 * 4 + 16);
  a.cand_idx = (int*)(base + 4096 * 4 + 16 + 8192 * 4);
  a.tokens = tokens; a.tok_slot_stride = tok_slot_stride; a.last_token = last_token; a.ctx = ctx;
  a.V = V; a.top_k = top_k; a.temperature = temperature; a.greedy = greedy; a.seed = seed;
  a.top_p = top_p; a.tok_ts = tok_ts; a.ts_slot_stride = ts_slot_stride; a.status = status;
  void* args[] = {&a};
  int rc = 0;
  if (!greedy) {
    rc = launch_small((const void*)sample_filter_kernel, dim3(64), dim3(256), args, use_pdl, stream);
    if (rc) return rc;
  }
  const size_t smem = (size_t)SF_CAND_MAX * 8;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(sample_final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    attr_set = true;
  }
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(1);
  cfg.blockDim = dim3(SMP_THREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = use_pdl ? 1 : 0;
  return (int)cudaLaunchKernelExC(&cfg, (const void*)sample_final_kernel, args);
}

int mdi_advance_step(int* ctx, int* state, int* pos_arr, int n_slots, int is_starter, int use_pdl, cudaStream_t stream) {
  void* args[] = {&ctx, &state, &pos_arr, &n_slots, &is_starter};
  return launch_small((const void*)advance_step_kernel, dim3(1), dim3(32), args, use_pdl, stream);
}

int mdi_wait_flag(const int* flag, const int* ctx, int* status, long long max_cycles, cudaStream_t stream) {
  wait_flag_kernel<<<1, 32, 0, stream>>>(flag, ctx, status, max_cycles);
  return (int)cudaGetLastError();
}
int mdi_set_flag(int* flag, const int* ctx, cudaStream_t stream) {
  set_flag_kernel<<<1, 32, 0, stream>>>(flag, ctx);
  return (int)cudaGetLastError();
}
int mdi_stamp(unsigned long long* dst, cudaStream_t stream) {
  stamp_kernel<<<1, 32, 0, stream>>>(dst);
  return (int)cudaGetLastError();
}
int mdi_copy_signal(const void* src, void* dst, size_t bytes, int* flag, unsigned int* done_ctr, const int* ctx,
                    const int* status, cudaStream_t stream) {
  if (bytes % 16) return -2;
  size_t nvec = bytes / 16;
  int blocks = (int)((nvec + 255) / 256);
  if (blocks > 296) blocks = 296;
  if (blocks < 1) blocks = 1;
  copy_signal_kernel<<<blocks, 256, 0, stream>>>((const uint4*)src, (uint4*)dst, nvec, HopSignal{flag, done_ctr, status}, ctx);
  return (int)cudaGetLastError();
}
int mdi_copy_bytes(const void* src, void* dst, size_t bytes, cudaStream_t stream) {
  if (bytes % 16) return -2;
  size_t nvec = bytes / 16;
  int blocks = (int)((nvec + 255) / 256);
  if (blocks > 1184) blocks = 1184;
  if (blocks < 1) blocks = 1;
  copy_vec_kernel<<<blocks, 256, 0, stream>>>((const uint4*)src, (uint4*)dst, nvec);
  return (int)cudaGetLastError();
}

}  // extern "C"

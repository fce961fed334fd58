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
__device__ __forceinline__ void stage_input(const bf16* __restrict__ x, const bf16* __restrict__ norm_w,
                                            float eps, int unit_offset, int K, bf16* xs, float* red) {
  const int tid = threadIdx.x;
  if (norm_w == nullptr) {
    const uint4* src = reinterpret_cast<const uint4*>(x);
    uint4* dst = reinterpret_cast<uint4*>(xs);
    for (int v = tid; v < K / 8; v += LIN_THREADS) dst[v] = __ldcg(src + v);
    __syncthreads();
    return;
  }
  float ss = 0.f;
  const uint4* src = reinterpret_cast<const uint4*>(x);
  uint4* dst = reinterpret_cast<uint4*>(xs);
  for (int v = tid; v < K / 8; v += LIN_THREADS) {
    uint4 r = __ldcg(src + v);
    dst[v] = r;
    float f;
    f = bf16lo(r.x); ss = fmaf(f, f, ss); f = bf16hi(r.x); ss = fmaf(f, f, ss);
    f = bf16lo(r.y); ss = fmaf(f, f, ss); f = bf16hi(r.y); ss = fmaf(f, f, ss);
    f = bf16lo(r.z); ss = fmaf(f, f, ss); f = bf16hi(r.z); ss = fmaf(f, f, ss);
    f = bf16lo(r.w); ss = fmaf(f, f, ss); f = bf16hi(r.w); ss = fmaf(f, f, ss);
  }
  ss = warp_sum(ss);
  if ((tid & 31) == 0) red[tid >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int w = 0; w < LIN_WARPS; ++w) tot += red[w];
  const float rstd = rsqrtf(tot / (float)K + eps);
  for (int i = tid; i < K; i += LIN_THREADS) {
    float xn = round_bf16(__bfloat162float(xs[i]) * rstd);
    float w = __bfloat162float(norm_w[i]);
    if (unit_offset) w = round_bf16(1.f + w);
    xs[i] = __float2bfloat16_rn(xn * w);
  }
  __syncthreads();
}

__device__ __forceinline__ float gelu_tanh(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  return 0.5f * x * (1.f + tanhf(k0 * (x + k1 * x * x * x)));
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float silu(float x) { return x / (1.f + __expf(-x)); }

enum Act { ACT_NONE = 0, ACT_SILU_GATE = 1, ACT_GELU_TANH_GATE = 2, ACT_GELU_ERF_GATE = 3, ACT_GELU_TANH = 4, ACT_GELU_ERF = 5 };

struct LinearArgs {
  const bf16* W;         // [N, K]
  const bf16* W2;        // gated: second projection [N, K]; else null
  const bf16* bias;      // [N] or null
  const bf16* bias2;     // [N] or null
  const bf16* x;         // input activation row(s): x + slot * x_slot_stride
  const bf16* norm_w;    // fused RMSNorm weight [K] or null
  const bf16* residual;  // residual + slot * res_slot_stride, [N], or null
  void* y;               // output (bf16, or fp32 when out_fp32): y + slot * y_slot_stride
  const int* ctx;
  long long x_slot_stride, res_slot_stride, y_slot_stride;  // in elements; 0 = not slotted
  int N, K;
  float eps;
  int unit_offset;
  int act;
  int out_fp32;
  int items_per_cta;
  HopWait wait;
  HopSignal signal;
};

__global__ void __launch_bounds__(LIN_THREADS) linear_decode_kernel(const LinearArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* xs = reinterpret_cast<bf16*>(smem_raw);
  __shared__ float red[LIN_WARPS];

  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bool gated = a.W2 != nullptr;
  const int n_items = gated ? a.N : (a.N + 1) / 2;
  const int item_lo = blockIdx.x * a.items_per_cta;
  const int item_hi = min(n_items, item_lo + a.items_per_cta);

  // Pull this CTA's first weight lines towards L2 while we (possibly) wait for the hop flag.
  if (item_lo < n_items) {
    const int r0 = gated ? item_lo : 2 * item_lo;
    const char* p = reinterpret_cast<const char*>(a.W + (size_t)r0 * a.K);
    const size_t bytes = (size_t)min(item_hi - item_lo, 4) * (gated ? 1 : 2) * a.K * sizeof(bf16);
    for (size_t off = (size_t)threadIdx.x * 128; off < bytes; off += (size_t)LIN_THREADS * 128) prefetch_l2(p + off);
  }
  pdl_wait_prior();
  hop_wait(a.wait, a.ctx);
  const int slot = a.ctx ? a.ctx[MDI_CTX_SLOT] : 0;
  const bf16* x = a.x + (size_t)slot * a.x_slot_stride;
  stage_input(x, a.norm_w, a.eps, a.unit_offset, a.K, xs, red);
  pdl_launch_dependents();

  const bf16* res = a.residual ? a.residual + (size_t)slot * a.res_slot_stride : nullptr;
  const uint4* xv = reinterpret_cast<const uint4*>(xs);
  const int nvec = a.K / 8;
  for (int it = item_lo + warp; it < item_hi; it += LIN_WARPS) {
    int ra, rb;
    const bf16 *wa, *wb;
    if (gated) {
      ra = rb = it;
      wa = a.W + (size_t)it * a.K;
      wb = a.W2 + (size_t)it * a.K;
    } else {
      ra = 2 * it;
      rb = min(2 * it + 1, a.N - 1);
      wa = a.W + (size_t)ra * a.K;
      wb = a.W + (size_t)rb * a.K;
    }
    float da, db;
    warp_dot2(wa, wb, xv, nvec, lane, da, db);
    if (lane == 0) {
      if (a.bias) {
        da += __bfloat162float(a.bias[ra]);
        if (!gated) db += __bfloat162float(a.bias[rb]);
      }
      if (gated && a.bias2) db += __bfloat162float(a.bias2[rb]);
      if (gated) {
        // reference rounds each projection to bf16, then act(a) (bf16) * b (bf16)
        float fa = round_bf16(da), fb = round_bf16(db), g;
        if (a.act == ACT_SILU_GATE) g = silu(fa);
        else if (a.act == ACT_GELU_TANH_GATE) g = gelu_tanh(fa);
        else g = gelu_erf(fa);
        float out = round_bf16(g) * fb;
        if (a.out_fp32) reinterpret_cast<float*>(a.y)[(size_t)slot * a.y_slot_stride + it] = out;
        else reinterpret_cast<bf16*>(a.y)[(size_t)slot * a.y_slot_stride + it] = __float2bfloat16_rn(out);
      } else {
        float o[2] = {da, db};
        int rows[2] = {ra, 2 * it + 1};
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          if (rows[j] >= a.N) continue;
          float v = o[j];
          if (a.act == ACT_GELU_TANH) v = gelu_tanh(round_bf16(v));
          else if (a.act == ACT_GELU_ERF) v = gelu_erf(round_bf16(v));
          if (res) v = round_bf16(v) + __bfloat162float(res[rows[j]]);
          if (a.out_fp32) reinterpret_cast<float*>(a.y)[(size_t)slot * a.y_slot_stride + rows[j]] = round_bf16(v);  // logits: bf16 values like nn.Linear, kept in fp32 for the sampler
          else reinterpret_cast<bf16*>(a.y)[(size_t)slot * a.y_slot_stride + rows[j]] = __float2bfloat16_rn(v);
        }
      }
    }
  }
  hop_signal(a.signal, a.ctx);
}

// ------------------------------------------------------------------------------------------------
// QKV projection with fused RoPE and KV-cache append.
// Weight rows are litGPT's group-interleaved layout: for each of G groups: q_per_kv query heads,
// one key head, one value head (model.py:686-699).  A warp owns the two rows RoPE mixes.
struct QKVArgs {
  const bf16* W;       // [(H + 2G) * hs, K]
  const bf16* bias;    // or null
  const bf16* x;       // + slot * x_slot_stride
  const bf16* norm_w;  // fused norm_1 weight or null
  const float* cos;    // [S, n_elem]
  const float* sin;
  bf16* q_out;         // [H * hs]
  bf16* kv;            // this layer's pool: [n_slots, 2, G, S, hs]
  const int* ctx;
  long long x_slot_stride;
  int K, n_head, n_groups, head_size, rope_n_elem, max_seq;
  float eps;
  int unit_offset;
  int items_per_cta;
  HopWait wait;
};

__global__ void __launch_bounds__(LIN_THREADS) qkv_decode_kernel(const QKVArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  bf16* xs = reinterpret_cast<bf16*>(smem_raw);
  __shared__ float red[LIN_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int hs = a.head_size, half_hs = hs / 2, ne = a.rope_n_elem, half_ne = ne / 2;
  const int qpk = a.n_head / a.n_groups;
  const int n_items = (a.n_head + 2 * a.n_groups) * half_hs;
  const int item_lo = blockIdx.x * a.items_per_cta;
  const int item_hi = min(n_items, item_lo + a.items_per_cta);

  pdl_wait_prior();
  hop_wait(a.wait, a.ctx);
  const int slot = a.ctx[MDI_CTX_SLOT], pos = a.ctx[MDI_CTX_POS];
  stage_input(a.x + (size_t)slot * a.x_slot_stride, a.norm_w, a.eps, a.unit_offset, a.K, xs, red);
  pdl_launch_dependents();

  const uint4* xv = reinterpret_cast<const uint4*>(xs);
  const int nvec = a.K / 8;
  for (int it = item_lo + warp; it < item_hi; it += LIN_WARPS) {
    const int j = it / half_hs, i = it % half_hs;  // head slot, pair index inside the head
    int ra, rb;
    const bool rot_pair = i < half_ne;
    if (rot_pair) { ra = i; rb = i + half_ne; }
    else { int t = i - half_ne; ra = ne + 2 * t; rb = ra + 1; }
    const size_t row0 = (size_t)j * hs;
    float da, db;
    warp_dot2(a.W + (row0 + ra) * a.K, a.W + (row0 + rb) * a.K, xv, nvec, lane, da, db);
    if (lane == 0) {
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
}

static inline int pick_items_per_cta(int n_items, int target_ctas) {
  // multiple of LIN_WARPS so every warp of a CTA gets the same number of row pairs
  int ipc = (n_items + target_ctas - 1) / target_ctas;
  ipc = ((ipc + LIN_WARPS - 1) / LIN_WARPS) * LIN_WARPS;
  return ipc < LIN_WARPS ? LIN_WARPS : ipc;
}

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

template <typename Args, typename Kern>
static int launch_pdl(Kern kern, const Args& args, int grid, size_t smem, cudaStream_t stream, int use_pdl) {
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

}  // namespace mdi

using namespace mdi;

extern "C" {

// ctas_per_sm controls the grid: grid ~= num_sms * ctas_per_sm row-pair chunks.
int mdi_linear_decode(const void* W, const void* W2, const void* bias, const void* bias2, const void* x,
                      const void* norm_w, const void* residual, void* y, const int* ctx, long long x_slot_stride,
                      long long res_slot_stride, long long y_slot_stride, int N, int K, float eps, int unit_offset,
                      int act, int out_fp32, const int* wait_flag, int* status, long long wait_max_cycles,
                      int* signal_flag, unsigned int* done_ctr, int ctas_per_sm, int use_pdl, cudaStream_t stream) {
  if (K % 8 != 0) return -2;
  LinearArgs a;
  a.W = (const bf16*)W; a.W2 = (const bf16*)W2; a.bias = (const bf16*)bias; a.bias2 = (const bf16*)bias2;
  a.x = (const bf16*)x; a.norm_w = (const bf16*)norm_w; a.residual = (const bf16*)residual; a.y = y; a.ctx = ctx;
  a.x_slot_stride = x_slot_stride; a.res_slot_stride = res_slot_stride; a.y_slot_stride = y_slot_stride;
  a.N = N; a.K = K; a.eps = eps; a.unit_offset = unit_offset; a.act = act; a.out_fp32 = out_fp32;
  a.wait = HopWait{wait_flag, status, wait_max_cycles};
  a.signal = HopSignal{signal_flag, done_ctr};
  const int n_items = W2 ? N : (N + 1) / 2;
  a.items_per_cta = pick_items_per_cta(n_items, num_sms() * (ctas_per_sm > 0 ? ctas_per_sm : 4));
  const int grid = (n_items + a.items_per_cta - 1) / a.items_per_cta;
  return launch_pdl(linear_decode_kernel, a, grid, (size_t)K * sizeof(bf16), stream, use_pdl);
}

int mdi_qkv_decode(const void* W, const void* bias, const void* x, const void* norm_w, const float* cos,
                   const float* sin, void* q_out, void* kv, const int* ctx, long long x_slot_stride, int K,
                   int n_head, int n_groups, int head_size, int rope_n_elem, int max_seq, float eps,
                   int unit_offset, const int* wait_flag, int* status, long long wait_max_cycles, int ctas_per_sm,
                   int use_pdl, cudaStream_t stream) {
  if (K % 8 != 0 || head_size % 2 != 0 || rope_n_elem % 2 != 0 || rope_n_elem > head_size) return -2;
  QKVArgs a;
  a.W = (const bf16*)W; a.bias = (const bf16*)bias; a.x = (const bf16*)x; a.norm_w = (const bf16*)norm_w;
  a.cos = cos; a.sin = sin; a.q_out = (bf16*)q_out; a.kv = (bf16*)kv; a.ctx = ctx; a.x_slot_stride = x_slot_stride;
  a.K = K; a.n_head = n_head; a.n_groups = n_groups; a.head_size = head_size; a.rope_n_elem = rope_n_elem;
  a.max_seq = max_seq; a.eps = eps; a.unit_offset = unit_offset;
  a.wait = HopWait{wait_flag, status, wait_max_cycles};
  const int n_items = (n_head + 2 * n_groups) * (head_size / 2);
  a.items_per_cta = pick_items_per_cta(n_items, num_sms() * (ctas_per_sm > 0 ? ctas_per_sm : 4));
  const int grid = (n_items + a.items_per_cta - 1) / a.items_per_cta;
  return launch_pdl(qkv_decode_kernel, a, grid, (size_t)K * sizeof(bf16), stream, use_pdl);
}

}  // extern "C"

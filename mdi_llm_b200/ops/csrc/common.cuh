// Shared device helpers for the sm_100a decode/prefill kernels of the B200 MDI engine.
//
// Conventions: activations and weights are bf16, all reductions/accumulators are fp32.
// "ctx" is the per-step device descriptor {slot, pos, wait, signal, token, step} written either by the host
// (host-driven scheduler, one tiny H2D per step) or by `mdi_advance_step` (device-driven
// pipeline, no host involvement) so that CUDA-graph replays never need new kernel arguments.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define MDI_CTX_SLOT 0
#define MDI_CTX_POS 1
#define MDI_CTX_WAIT 2    // value the incoming-hop flag of `slot` must reach before reading input
#define MDI_CTX_SIGNAL 3  // value published to the next stage's flag of `slot` after the output
#define MDI_CTX_TOKEN 4   // input token id of this step (starter)
#define MDI_CTX_STEP 5    // global step counter (device-driven mode)
#define MDI_CTX_INTS 8

#define MDI_CHECK(call)                                            \
  do {                                                             \
    cudaError_t _e = (call);                                       \
    if (_e != cudaSuccess) return (int)_e;                         \
  } while (0)

namespace mdi {

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// two packed bf16 -> two fp32 (exact: bf16 is the top half of an fp32)
__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ float round_bf16(float f) { return __bfloat162float(__float2bfloat16_rn(f)); }

// streaming 16-byte weight load (evict-first).  Deliberately an intrinsic, not `asm volatile`:
// volatile asm pins the load next to its use and ptxas then keeps only 2 loads in flight.
__device__ __forceinline__ uint4 ldg_stream(const void* p) { return __ldcs(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// ---- cross-GPU flag protocol (system scope: producer and consumer are different devices) ----
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_volatile(const int* p) {
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Bounded spin on a hop flag.  Returns false on watchdog expiry (the caller records an error
// code in the status word instead of hanging the GPU — SURVEY §5.3).
__device__ __forceinline__ bool wait_flag_ge(const int* flag, int want, long long max_cycles) {
  long long t0 = clock64();
  while (ld_acquire_sys(flag) < want) {
    __nanosleep(64);
    if (max_cycles > 0 && clock64() - t0 > max_cycles) return false;
  }
  return true;
}

// ---- abort propagation ---------------------------------------------------------------------------------
// Flags carry round numbers (small positive ints).  MDI_POISON is larger than any of them, so a poisoned
// flag satisfies every wait at once.  A stage that (a) trips its watchdog or (b) reads a poisoned flag marks
// itself aborted (status[1] = 1); an aborted stage never waits again and publishes MDI_POISON instead of its
// round number, so the abort travels round the ring at hop speed and every rank's queued steps drain in
// microseconds instead of each one spinning out its own watchdog budget.  The host can start an abort by
// writing MDI_POISON into a node's own flags (DevicePipeline.poison / PUT /stop).
// status words: [0] error bits (1 = hop watchdog expired here, 2 = intra-stage dependency watchdog,
//               4 = sampler candidate overflow), [1] aborted, [2..3] cycles CTA 0 spent waiting (64-bit).
#define MDI_POISON 0x7fffffff

struct HopWait {   // consumer side; flag == nullptr disables
  const int* flag;       // base of this GPU's flag array (one int per sample slot)
  int* status;           // status words (see above), may be null
  long long max_cycles;  // 0 = wait forever
};
struct HopSignal {  // producer side; flag == nullptr disables
  int* flag;               // base of the NEXT stage's flag array (peer-mapped or local)
  unsigned int* done_ctr;  // local counter of finished CTAs (self-resetting)
  const int* status;       // this stage's status words (abort -> publish poison), may be null
};

// one thread: wait until flag >= want, honouring / recording aborts
__device__ __forceinline__ void hop_wait_one(const int* flag, int want, int* status, long long max_cycles) {
  if (status && ld_volatile(status + 1) != 0) return;  // already aborted: drain
  const long long t0 = clock64();
  int v;
  while ((v = ld_acquire_sys(flag)) < want) {
    __nanosleep(64);
    if (max_cycles > 0 && clock64() - t0 > max_cycles) {
      if (status) { atomicOr(status, 1); atomicExch(status + 1, 1); }
      return;
    }
  }
  if (v == MDI_POISON && status) atomicExch(status + 1, 1);
}

// All threads call; thread 0 spins, everybody leaves after the flag for ctx's slot >= ctx's wait value.
// status[2..3] (as one 64-bit word) accumulates the cycles CTA 0 spent waiting — the *exposed* hop/idle
// time the benchmark reports per token.
__device__ __forceinline__ void hop_wait(const HopWait& w, const int* ctx) {
  if (w.flag == nullptr) return;
  if (threadIdx.x == 0) {
    const int slot = ctx[MDI_CTX_SLOT], want = ctx[MDI_CTX_WAIT];
    const long long t0 = clock64();
    hop_wait_one(w.flag + slot, want, w.status, w.max_cycles);
    if (w.status && blockIdx.x == 0 && blockIdx.y == 0)
      atomicAdd(reinterpret_cast<unsigned long long*>(w.status + 2), (unsigned long long)(clock64() - t0));
  }
  __syncthreads();
}

// Called by every CTA after its output stores.  The last CTA to arrive publishes the flag.
// Ordering: the CTA's stores -> bar.sync -> thread 0's system-scope fence (cumulative) -> ticket;
// the last ticket holder fences again and releases the flag (same pattern as a grid barrier).
__device__ __forceinline__ void hop_signal(const HopSignal& s, const int* ctx) {
  if (s.flag == nullptr) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
    const unsigned int prev = atomicAdd(s.done_ctr, 1u);
    if (prev == total - 1) {
      *s.done_ctr = 0;  // ready for the next launch (stream-ordered)
      __threadfence_system();
      const bool aborted = s.status != nullptr && ld_volatile(s.status + 1) != 0;
      st_release_sys(s.flag + ctx[MDI_CTX_SLOT], aborted ? MDI_POISON : ctx[MDI_CTX_SIGNAL]);
    }
  }
}

// Decode hop, cheap form: every CTA has written its part of the output row to LOCAL memory; tickets are taken
// at gpu scope (fence + atomic, the classic "last block" pattern) and only the last CTA touches the peer: it
// copies the finished row over NVLink with 16-byte stores, fences once at system scope and releases the flag.
// (hop_signal above makes EVERY CTA fence at system scope after its few remote bytes: measured +8-10 us on the
// tail of the stage's last kernel, because a MEMBAR.SYS waits behind the SM's in-flight weight stream.)
// `pre` (n_pre elements, may be null): a second local row that goes in FRONT of the main one — the residual
// stream x of an [x | h] message (stage boundary between a gated MLP's gate/up and down projections).
__device__ __forceinline__ void hop_signal_copy(const HopSignal& s, const int* ctx, const __nv_bfloat16* local_row,
                                                __nv_bfloat16* remote_row, int n_elems,
                                                const __nv_bfloat16* pre = nullptr, int n_pre = 0) {
  if (s.flag == nullptr) return;
  __shared__ int hop_is_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
    const unsigned int prev = atomicAdd(s.done_ctr, 1u);
    hop_is_last = (prev == total - 1);
    if (hop_is_last) *s.done_ctr = 0;
  }
  __syncthreads();
  if (!hop_is_last) return;
  __threadfence();
  if (pre != nullptr && n_pre > 0) {
    const uint4* psrc = reinterpret_cast<const uint4*>(pre);
    uint4* pdst = reinterpret_cast<uint4*>(remote_row);
    for (int v = threadIdx.x; v < n_pre / 8; v += blockDim.x) pdst[v] = __ldcg(psrc + v);
    remote_row += n_pre;
  }
  const uint4* src = reinterpret_cast<const uint4*>(local_row);
  uint4* dst = reinterpret_cast<uint4*>(remote_row);
  for (int v = threadIdx.x; v < n_elems / 8; v += blockDim.x) dst[v] = __ldcg(src + v);
  for (int i = (n_elems / 8) * 8 + threadIdx.x; i < n_elems; i += blockDim.x) remote_row[i] = local_row[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    const bool aborted = s.status != nullptr && ld_volatile(s.status + 1) != 0;
    st_release_sys(s.flag + ctx[MDI_CTX_SLOT], aborted ? MDI_POISON : ctx[MDI_CTX_SIGNAL]);
  }
}

// ---- intra-stage dependencies by flag instead of grid completion -----------------------------------------
// A PDL-launched consumer is already resident while its producer runs; `griddepcontrol.wait` releases it
// only after the producer grid has completed AND flushed (~1.4 us after the last CTA's exit, measured).
// With these flags the producer's last CTA (ticket) publishes `flag = step + 1` right after its stores
// and the consumer's thread 0 acquires it (gpu scope; the acquire invalidates the SM's L1), which cuts
// the boundary to the ticket + poll latency.  Legal because PDL starts a dependent grid only once every
// producer CTA is resident (no deadlock) and every kernel of a stage depends on its predecessor (the
// chain is linear, so WAR hazards on the ping-pong buffers are covered transitively).
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
struct DepWait {   // flag == nullptr: fall back to griddepcontrol.wait
  const int* flag;
  int* status;           // watchdog error word (set to 2 on expiry), may be null
  long long max_cycles;  // 0 = wait forever
};
struct DepSignal {  // flag == nullptr disables
  int* flag;
  unsigned int* ctr;  // ticket counter (self-resetting)
};
// all threads call; returns after the producer's flag shows this step
__device__ __forceinline__ void dep_wait(const DepWait& w, const int* ctx) {
  if (threadIdx.x == 0) {
    const int want = ctx[MDI_CTX_STEP] + 1;
    const long long t0 = clock64();
    // a producer is one kernel of this stage (< 1 ms); cap the watchdog at 2e8 cycles so that a protocol bug
    // shows up as status 2 within a fraction of a second per launch instead of hanging the GPU
    const long long cap = (w.max_cycles > 0 && w.max_cycles < 200000000ll) ? w.max_cycles : 200000000ll;
    // poll with plain volatile loads (an acquire per poll would invalidate this SM's L1 every iteration and
    // disturb the producer CTAs that share it); one acquire when the value is there orders the data reads
    while (ld_volatile(w.flag) < want) {
      __nanosleep(20);
      if (clock64() - t0 > cap) {
        if (w.status) atomicExch(w.status, 2);
        break;
      }
    }
    (void)ld_acquire_gpu(w.flag);
  }
  __syncthreads();
}
// every CTA calls after its last output store (idle CTAs too: the ticket count is the whole grid).
// Same-address atomics from hundreds of CTAs serialise in L2 (~14 ns each), so the tickets are two-level:
// 16 sub-counters on separate 128-byte lines (CTA index mod 16), whose last arrivals meet on a top counter.
constexpr int DEP_FAN = 16;
constexpr int DEP_CTR_STRIDE = 32;  // uints: one L2 line per counter; ctr[0] = top, ctr[(1 + i) * 32] = sub i
__device__ __forceinline__ void dep_signal(const DepSignal& s, const int* ctx) {
  if (s.flag == nullptr) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
    const unsigned int me = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    const unsigned int fan = total < (unsigned)DEP_FAN ? total : (unsigned)DEP_FAN;
    const unsigned int sub = me % fan;
    const unsigned int n_sub = (total - sub + fan - 1) / fan;  // CTAs that share this sub-counter
    unsigned int* c_sub = s.ctr + (1 + sub) * DEP_CTR_STRIDE;
    if (atomicAdd(c_sub, 1u) == n_sub - 1) {
      *c_sub = 0;
      __threadfence();
      if (atomicAdd(s.ctr, 1u) == fan - 1) {
        *s.ctr = 0;
        __threadfence();
        st_release_gpu(s.flag, ctx[MDI_CTX_STEP] + 1);
      }
    }
  }
}

// ---- mbarrier + 1-D bulk async copy (TMA engine, no tensor map needed) ------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy, completion counted in bytes on `bar`; L2 evict-first hint for weights
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- activations (shared by the decode linears and the tcgen05 GEMM epilogue) ---------------------------
__device__ __forceinline__ float gelu_tanh(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  return 0.5f * x * (1.f + tanhf(k0 * (x + k1 * x * x * x)));
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float silu(float x) { return x / (1.f + __expf(-x)); }

enum Act { ACT_NONE = 0, ACT_SILU_GATE = 1, ACT_GELU_TANH_GATE = 2, ACT_GELU_ERF_GATE = 3, ACT_GELU_TANH = 4, ACT_GELU_ERF = 5 };

// gate activation of a gated MLP (ACT_*_GATE)
__device__ __forceinline__ float apply_act(float x, int act) {
  if (act == ACT_SILU_GATE) return silu(x);
  if (act == ACT_GELU_TANH_GATE) return gelu_tanh(x);
  return gelu_erf(x);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// L2 prefetch through the TMA engine (no shared-memory destination): extends the look-ahead of a
// weight stream beyond what the shared-memory ring can hold
__device__ __forceinline__ void bulk_prefetch_l2(const void* gsrc, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gsrc), "r"(bytes) : "memory");
}

// ---- device-side tracer ---------------------------------------------------------------------------
// Each traced kernel owns one record of 8 x u64 (ns, %globaltimer):
//   [0] min entry  [1] min "input ready" (after PDL/hop wait)  [2] max "staged"  [3] min exit  [4] max exit  [5] #CTAs
//   [6] optional device pointer to a per-CTA table (8 x u64 per CTA: the five marks, [5] = SM id)  [7] its capacity (CTAs)
// Recording is per CTA (thread 0), costs a handful of atomics and is off when `rec == nullptr`.
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void trace_mark(unsigned long long* rec, int field, bool is_min) {
  if (rec == nullptr || threadIdx.x != 0) return;
  const unsigned long long t = globaltimer_ns();
  if (is_min) atomicMin(rec + field, t); else atomicMax(rec + field, t);
  if (field == 4) atomicAdd(rec + 5, 1ull);
  unsigned long long* detail = reinterpret_cast<unsigned long long*>(rec[6]);
  if (detail != nullptr && blockIdx.x + gridDim.x * blockIdx.y < rec[7]) {
    unsigned long long* d = detail + (size_t)(blockIdx.x + gridDim.x * blockIdx.y) * 8;
    d[field] = t;
    if (field == 0) {
      unsigned int smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      d[5] = smid;
    }
  }
}

// extra per-CTA phase marks (detail table fields 6 and 7), only recorded when a detail table is attached
__device__ __forceinline__ void trace_phase(unsigned long long* rec, int field) {
  if (rec == nullptr || threadIdx.x != 0) return;
  unsigned long long* detail = reinterpret_cast<unsigned long long*>(rec[6]);
  if (detail != nullptr && blockIdx.x + gridDim.x * blockIdx.y < rec[7])
    detail[(size_t)(blockIdx.x + gridDim.x * blockIdx.y) * 8 + field] = globaltimer_ns();
}

// Programmatic dependent launch: let the next kernel's prologue overlap our tail, and wait for
// the previous kernel's memory before touching activations.
__device__ __forceinline__ void pdl_wait_prior() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void pdl_launch_dependents() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// host: SM count of the current device
inline int device_sm_count() {
  int dev = 0, n = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n > 0 ? n : 148;
}

}  // namespace mdi

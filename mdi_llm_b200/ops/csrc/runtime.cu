// Native runtime pieces around the kernels: peer-memory (CUDA IPC) buffers for the fused hop,
// CUDA-graph capture/replay of a stage's decode step, device-side timing helpers.
//
// The reference's data plane is sockets + pickle driven by Python threads
// (src/sub/connections.py); here the only host work per decode step is one cudaGraphLaunch —
// issued from this C loop so that the Python interpreter is not in the per-token path at all.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#define RT_CHECK(call)                      \
  do {                                      \
    cudaError_t _e = (call);                \
    if (_e != cudaSuccess) return (int)_e;  \
  } while (0)

extern "C" {

// sources digest this binary was built from (ops/build.py compares it with the work tree before loading)
#ifndef MDI_BUILD_DIGEST_STR
#define MDI_BUILD_DIGEST_STR "unknown"
#endif
const char* mdi_build_digest() { return "MDI_BUILD_DIGEST:" MDI_BUILD_DIGEST_STR; }

const char* mdi_error_string(int code) { return cudaGetErrorString((cudaError_t)code); }

int mdi_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* total_mem) {
  int dev = 0;
  RT_CHECK(cudaGetDevice(&dev));
  cudaDeviceProp p;
  RT_CHECK(cudaGetDeviceProperties(&p, dev));
  *sm_count = p.multiProcessorCount;
  *cc_major = p.major;
  *cc_minor = p.minor;
  *total_mem = p.totalGlobalMem;
  return 0;
}

// ---- peer memory ------------------------------------------------------------------------------
// A hop buffer is plain cudaMalloc memory (IPC handles cannot be taken on VMM/expandable
// segments of the torch allocator), zero-initialised, exported as a 64-byte handle that the next
// stage's process opens.  Opening enables peer access lazily.
int mdi_p2p_alloc(size_t bytes, void** ptr, unsigned char* handle64) {
  RT_CHECK(cudaMalloc(ptr, bytes));
  RT_CHECK(cudaMemset(*ptr, 0, bytes));
  RT_CHECK(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  RT_CHECK(cudaIpcGetMemHandle(&h, *ptr));
  memcpy(handle64, &h, sizeof(h));
  return 0;
}
int mdi_p2p_open(const unsigned char* handle64, void** ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  RT_CHECK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
int mdi_p2p_close(void* ptr) { return (int)cudaIpcCloseMemHandle(ptr); }
int mdi_p2p_free(void* ptr) { return (int)cudaFree(ptr); }

// same-process multi-GPU (tests, single-process launcher): direct peer access
int mdi_enable_peer(int dev, int peer) {
  int can = 0;
  RT_CHECK(cudaDeviceCanAccessPeer(&can, dev, peer));
  if (!can) return -4;
  int cur = 0;
  RT_CHECK(cudaGetDevice(&cur));
  RT_CHECK(cudaSetDevice(dev));
  cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
  cudaSetDevice(cur);
  return (int)e;
}

// ---- CUDA graphs --------------------------------------------------------------------------------
int mdi_graph_begin(cudaStream_t stream) {
  return (int)cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal);
}
int mdi_graph_end(cudaStream_t stream, cudaGraphExec_t* exec, int* n_nodes) {
  cudaGraph_t g = nullptr;
  RT_CHECK(cudaStreamEndCapture(stream, &g));
  size_t n = 0;
  RT_CHECK(cudaGraphGetNodes(g, nullptr, &n));
  if (n_nodes) *n_nodes = (int)n;
  RT_CHECK(cudaGraphInstantiate(exec, g, 0));
  RT_CHECK(cudaGraphDestroy(g));
  return 0;
}
int mdi_graph_launch(cudaGraphExec_t exec, cudaStream_t stream, int times) {
  for (int i = 0; i < times; ++i) RT_CHECK(cudaGraphLaunch(exec, stream));
  return 0;
}
int mdi_graph_destroy(cudaGraphExec_t exec) { return (int)cudaGraphExecDestroy(exec); }

// Two graphs alternated (starter: full step graph, then per-round bookkeeping), kept in C so the
// launch loop of a whole generation is one foreign call.
int mdi_graph_launch_pattern(cudaGraphExec_t a, int times_a, cudaGraphExec_t b, int times_b, int repeats,
                             cudaStream_t stream) {
  for (int r = 0; r < repeats; ++r) {
    for (int i = 0; i < times_a; ++i) RT_CHECK(cudaGraphLaunch(a, stream));
    for (int i = 0; i < times_b; ++i) RT_CHECK(cudaGraphLaunch(b, stream));
  }
  return 0;
}

// ---- pinned host memory + async copies (e2e path: step inputs H2D, tokens D2H) ------------------
int mdi_host_alloc(size_t bytes, void** ptr) { return (int)cudaHostAlloc(ptr, bytes, cudaHostAllocDefault); }
int mdi_host_free(void* ptr) { return (int)cudaFreeHost(ptr); }
int mdi_memcpy_async(void* dst, const void* src, size_t bytes, int kind, cudaStream_t stream) {
  return (int)cudaMemcpyAsync(dst, src, bytes, (cudaMemcpyKind)kind, stream);
}
int mdi_stream_sync(cudaStream_t stream) { return (int)cudaStreamSynchronize(stream); }

}  // extern "C"

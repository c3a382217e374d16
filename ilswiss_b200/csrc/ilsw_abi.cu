// ilswiss_b200 -- C ABI (include/ilswiss_b200.h) over the sm_100a kernels.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <string>
#include <vector>

#include "../../include/ilswiss_b200.h"
#include "ilsw_engine.cuh"
#include "ilsw_tmap.h"
#include "ilsw_program.h"

using namespace ilsw;

// ------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CU(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e__ = (call);                                                                     \
    if (e__ != cudaSuccess)                                                                       \
      return fail(ILSW_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

extern "C" int ilsw_abi_version(void) { return ILSW_ABI_VERSION; }
extern "C" const char* ilsw_last_error(void) { return g_err; }

extern "C" int ilsw_device_info(int* sm_count, int* cc_major, int* cc_minor, char* name, int name_len) {
  int dev = 0;
  CU(cudaGetDevice(&dev));
  cudaDeviceProp p;
  CU(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  if (name && name_len > 0) snprintf(name, name_len, "%s", p.name);
  return ILSW_OK;
}

extern "C" int ilsw_mlp_num_params(int in_dim, int hidden, int out_dim, int log_std_head) {
  return mlp_num_params(in_dim, hidden, out_dim, log_std_head);
}

// ==========================================================================================
// Replay ring (R1-R4)
// ==========================================================================================
struct ilsw_rb {
  int64_t capacity;
  int O, A, stride, host_w;
  float* rows;      // [capacity x stride]
  float* cold;      // [capacity x 4]
  int64_t top, size;
  float* staging;   // device staging [staging_cap x host_w]
  int64_t staging_cap, pending;
  cudaEvent_t staged;
  bool staged_valid;
  cudaStream_t last_stream;   // compute stream of the latest commit (scatter kernels run there, never on the copy stream)
};

// R2 commit: staged transitions -> hot rows at (top + i) % capacity (+ cold rows).  The destination is a CONTIGUOUS range of
// ring rows (at most one wrap), so the kernel is a flat, fully coalesced copy over the destination floats: thread -> output
// float e = row * stride + k, 4 independent elements per thread in flight; the source is the packed staging row (host_w
// floats, 4-byte aligned).  (Round 1: one warp per row, one dependent load -> store and a 64-bit modulo per row: 0.12-0.35 of
// the HBM roofline, profiles/r1_replay_bench.txt.)
__global__ void __launch_bounds__(256) rb_scatter_kernel(const float* __restrict__ staging, int64_t n, float* rows,
                                                          float* cold, int64_t top, int64_t capacity, int O, int A,
                                                          int stride, int host_w) {
  const int hot = 2 * O + A + 2;
  const int64_t total = n * stride;
  constexpr int U = 4;
  const int64_t e0 = ((int64_t)blockIdx.x * blockDim.x * U) + threadIdx.x;
  float v[U];
  int64_t slot[U]; int kk[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const int64_t e = e0 + (int64_t)u * blockDim.x;
    v[u] = 0.f; slot[u] = -1; kk[u] = 0;
    if (e < total) {
      const int64_t i = e / stride;
      const int k = (int)(e - i * stride);
      int64_t sl = top + i;
      if (sl >= capacity) sl -= capacity;           // i < capacity (commits are chunked): one wrap at most
      slot[u] = sl; kk[u] = k;
      if (k < hot) v[u] = __ldg(staging + i * host_w + k);
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (slot[u] >= 0) rows[slot[u] * stride + kk[u]] = v[u];
  // cold rows: 4 floats per transition (absorbing flags / timeout), one thread per float
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n * 4) {
    const int64_t i = c >> 2; const int l = (int)(c & 3);
    int64_t sl = top + i;
    if (sl >= capacity) sl -= capacity;
    cold[sl * 4 + l] = l < 3 ? __ldg(staging + i * host_w + hot + l) : 0.f;
  }
}

// R3/R4: uniform sample + minibatch gather.  A group of LPR lanes (8 / 16 / 32, chosen from the row length) owns one sampled
// row; rows are 16-byte aligned (stride % 4 == 0), so every access is a 128-bit ld.global.nc / st.global and the lanes
// of a group read consecutive 16-byte vectors of the same row (coalesced 128..512-byte segments -- the rows themselves
// are random HBM addresses).  Rows of 17..64 vectors (Ant) keep two loads per lane in flight before the first store.  Optional in-kernel Philox index (speed mode).
template <int LPR, int kGatherUnroll>
__global__ void __launch_bounds__(256, 8) rb_gather_kernel(const float* __restrict__ rows, const float* __restrict__ cold,
                                                         const int32_t* __restrict__ idx, int B, int stride,
                                                         float* out_hot, float* out_cold, int64_t size, uint64_t seed,
                                                         uint64_t counter, int32_t* idx_out) {
  constexpr int kRowsPerCta = 256 / LPR;
  const int sub = threadIdx.x % LPR;
  const int b = blockIdx.x * kRowsPerCta + threadIdx.x / LPR;
  if (b >= B) return;
  int64_t r;
  if (idx) r = idx[b];
  else {
    r = philox_index(seed, (uint32_t)counter, (uint32_t)b, (uint32_t)(counter >> 32) + 0x51u, (int)size);
    if (idx_out && sub == 0) idx_out[b] = (int32_t)r;
  }
  const float4* src = reinterpret_cast<const float4*>(rows + r * stride);
  float4* dst = reinterpret_cast<float4*>(out_hot + (size_t)b * stride);
  const int nv = stride >> 2;
  for (int k0 = sub; k0 < nv; k0 += LPR * kGatherUnroll) {
    float4 v[kGatherUnroll];
#pragma unroll
    for (int u = 0; u < kGatherUnroll; ++u)
      if (k0 + u * LPR < nv) v[u] = __ldg(src + k0 + u * LPR);
#pragma unroll
    for (int u = 0; u < kGatherUnroll; ++u)
      if (k0 + u * LPR < nv) dst[k0 + u * LPR] = v[u];
  }
  if (out_cold && sub == 0)
    reinterpret_cast<float4*>(out_cold)[b] = __ldg(reinterpret_cast<const float4*>(cold) + r);
}

static void launch_gather(const ilsw_rb* rb, const int32_t* idx, int B, float* out_hot, float* out_cold, uint64_t seed,
                          uint64_t counter, int32_t* idx_out, cudaStream_t st);

extern "C" int ilsw_rb_create(ilsw_rb** out, int64_t capacity, int obs_dim, int act_dim) {
  if (!out || capacity <= 0 || obs_dim <= 0 || act_dim <= 0) return fail(ILSW_ERR_ARG, "rb_create: bad arguments");
  if (capacity > 0x7fffffffLL) return fail(ILSW_ERR_ARG, "rb_create: capacity must fit int32 indices");
  ilsw_rb* rb = new ilsw_rb();
  memset(rb, 0, sizeof(*rb));
  rb->capacity = capacity; rb->O = obs_dim; rb->A = act_dim;
  // rows start on 64-byte boundaries (the DRAM access granule): a sampled row then costs exactly its own blocks -- with the
  // packed 16-byte-aligned layout of round 1 a 112-byte Hopper row dragged in 1.69x its size (profiles/r1_replay_bench.txt)
  rb->stride = round_up(2 * obs_dim + act_dim + 2, 16);
  rb->host_w = 2 * obs_dim + act_dim + 5;
  cudaError_t e = cudaMalloc(&rb->rows, (size_t)capacity * rb->stride * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&rb->cold, (size_t)capacity * 4 * sizeof(float));
  if (e == cudaSuccess) e = cudaMemset(rb->cold, 0, (size_t)capacity * 4 * sizeof(float));
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&rb->staged, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    int code = fail(ILSW_ERR_CUDA, "rb_create: %s (capacity %lld x %d floats)", cudaGetErrorString(e), (long long)capacity, rb->stride);
    if (rb->rows) cudaFree(rb->rows);
    if (rb->cold) cudaFree(rb->cold);
    delete rb;
    return code;
  }
  *out = rb;
  return ILSW_OK;
}
extern "C" int ilsw_rb_destroy(ilsw_rb* rb) {
  if (!rb) return ILSW_OK;
  cudaFree(rb->rows); cudaFree(rb->cold);
  if (rb->staging) cudaFree(rb->staging);
  cudaEventDestroy(rb->staged);
  delete rb;
  return ILSW_OK;
}
extern "C" int ilsw_rb_host_row_floats(const ilsw_rb* rb) { return rb ? rb->host_w : ILSW_ERR_ARG; }
extern "C" int ilsw_rb_row_stride(const ilsw_rb* rb) { return rb ? rb->stride : ILSW_ERR_ARG; }
extern "C" int64_t ilsw_rb_capacity(const ilsw_rb* rb) { return rb ? rb->capacity : ILSW_ERR_ARG; }
extern "C" int64_t ilsw_rb_size(const ilsw_rb* rb) { return rb ? rb->size + 0 : ILSW_ERR_ARG; }
extern "C" int64_t ilsw_rb_top(const ilsw_rb* rb) { return rb ? rb->top : ILSW_ERR_ARG; }
extern "C" float* ilsw_rb_rows_ptr(ilsw_rb* rb) { return rb ? rb->rows : nullptr; }
extern "C" int ilsw_rb_clear(ilsw_rb* rb) {
  if (!rb) return fail(ILSW_ERR_ARG, "rb_clear: null");
  rb->top = rb->size = rb->pending = 0;
  return ILSW_OK;
}

extern "C" int ilsw_rb_set_cursor(ilsw_rb* rb, int64_t top, int64_t size) {
  if (!rb || top < 0 || top >= rb->capacity || size < 0 || size > rb->capacity) return fail(ILSW_ERR_ARG, "rb_set_cursor: bad arguments");
  if (rb->pending) return fail(ILSW_ERR_STATE, "rb_set_cursor: pending appends");
  rb->top = top; rb->size = size;
  return ILSW_OK;
}

extern "C" int ilsw_rb_append(ilsw_rb* rb, const float* host_rows, int64_t n, void* copy_stream) {
  if (!rb || (!host_rows && n > 0) || n < 0) return fail(ILSW_ERR_ARG, "rb_append: bad arguments");
  if (n == 0) return ILSW_OK;
  if (n > rb->capacity) return fail(ILSW_ERR_ARG, "rb_append: burst larger than the ring");
  cudaStream_t cs = (cudaStream_t)copy_stream;
  // more pending rows than ring slots would make two rows of one scatter race for a slot: drain first.  The scatter
  // always runs on the COMPUTE stream last used with this ring (never on the copy stream, where it could overwrite rows
  // an engine kernel is gathering); the copy stream then waits for it before reusing the staging area.
  if (rb->pending > 0 && (rb->pending + n > rb->capacity || rb->pending + n > rb->staging_cap)) {
    int rc = ilsw_rb_commit(rb, (void*)rb->last_stream);
    if (rc) return rc;
  }
  if (rb->pending == 0 && rb->staged_valid) CU(cudaStreamWaitEvent(cs, rb->staged, 0));  // staging reuse after the scatter
  if (n > rb->staging_cap) {
    int64_t cap = rb->staging_cap ? rb->staging_cap : 1024;
    while (cap < n) cap *= 2;
    if (rb->staging) { CU(cudaDeviceSynchronize()); CU(cudaFree(rb->staging)); }
    CU(cudaMalloc(&rb->staging, (size_t)cap * rb->host_w * sizeof(float)));
    rb->staging_cap = cap;
  }
  CU(cudaMemcpyAsync(rb->staging + rb->pending * rb->host_w, host_rows, (size_t)n * rb->host_w * sizeof(float),
                     cudaMemcpyHostToDevice, cs));
  CU(cudaEventRecord(rb->staged, cs));
  rb->staged_valid = true;
  rb->pending += n;
  return ILSW_OK;
}

extern "C" int ilsw_rb_commit(ilsw_rb* rb, void* stream) {
  if (!rb) return fail(ILSW_ERR_ARG, "rb_commit: null");
  if (rb->pending == 0) return ILSW_OK;
  cudaStream_t st = (cudaStream_t)stream;
  rb->last_stream = st;
  if (rb->staged_valid) CU(cudaStreamWaitEvent(st, rb->staged, 0));
  const int64_t n = rb->pending;
  const int64_t per_cta = 256 * 4;                 // destination floats per CTA (rb_scatter_kernel: 4 per thread)
  rb_scatter_kernel<<<(unsigned)((n * rb->stride + per_cta - 1) / per_cta), 256, 0, st>>>(rb->staging, n, rb->rows, rb->cold, rb->top,
                                                                                        rb->capacity, rb->O, rb->A, rb->stride, rb->host_w);
  CU(cudaGetLastError());
  // the staging area may be overwritten by the next append on the copy stream only after this
  // scatter has run: make later copies wait on it
  CU(cudaEventRecord(rb->staged, st));
  rb->top = (rb->top + n) % rb->capacity;
  rb->size = rb->size + n > rb->capacity ? rb->capacity : rb->size + n;
  rb->pending = 0;
  return ILSW_OK;
}

extern "C" int ilsw_rb_load_device(ilsw_rb* rb, const float* dev_hot_rows, int64_t n, void* stream) {
  if (!rb || !dev_hot_rows || n < 0 || n > rb->capacity) return fail(ILSW_ERR_ARG, "rb_load_device: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = ilsw_rb_commit(rb, stream);
  if (rc) return rc;
  const size_t rowb = (size_t)rb->stride * sizeof(float);
  int64_t first = n < rb->capacity - rb->top ? n : rb->capacity - rb->top;
  CU(cudaMemcpyAsync(rb->rows + rb->top * rb->stride, dev_hot_rows, (size_t)first * rowb, cudaMemcpyDeviceToDevice, st));
  CU(cudaMemsetAsync(rb->cold + rb->top * 4, 0, (size_t)first * 16, st));
  if (n > first) {
    CU(cudaMemcpyAsync(rb->rows, dev_hot_rows + first * rb->stride, (size_t)(n - first) * rowb, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemsetAsync(rb->cold, 0, (size_t)(n - first) * 16, st));
  }
  rb->top = (rb->top + n) % rb->capacity;
  rb->size = rb->size + n > rb->capacity ? rb->capacity : rb->size + n;
  return ILSW_OK;
}

// R3/R4 on the TMA unit: ring rows are 64-byte aligned (stride % 16 floats == 0, see ilsw_rb_create), so a sampled row is ONE
// 1-D bulk copy (cp.async.bulk global -> shared, SASS UBLKCP) issued by one lane; a warp stages up to kGatherStageBytes
// of rows behind its own mbarrier and, because consecutive batch rows are contiguous in the output tile, writes them back
// with ONE bulk store (shared -> global).  No data passes through registers; only the sampled indices (Philox or caller
// supplied) and the 16-byte cold rows do.  DRAM sees whole 64-byte blocks of exactly the sampled rows.
constexpr int kGatherStageBytes = 4096;
__global__ void __launch_bounds__(256) rb_gather_bulk_kernel(const float* __restrict__ rows, const float* __restrict__ cold,
                                                             const int32_t* __restrict__ idx, int B, int stride, float* out_hot,
                                                             float* out_cold, int64_t size, uint64_t seed, uint64_t counter,
                                                             int32_t* idx_out, int rows_per_warp) {
  extern __shared__ __align__(128) unsigned char gsm[];
  __shared__ unsigned long long bars[8];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t row_bytes = (uint32_t)stride * 4u;
  const uint32_t stage_bytes = (uint32_t)rows_per_warp * row_bytes;
  const uint32_t sm_w = tc5::smem_u32(gsm) + (uint32_t)w * stage_bytes;
  if (lane == 0) tc5::mbar_init(&bars[w], 1);
  tc5::fence_barrier_init();
  __syncthreads();
  uint32_t phase = 0;
  const int64_t warps_total = (int64_t)gridDim.x * 8;
  const int64_t step = warps_total * rows_per_warp;
  int64_t base = ((int64_t)blockIdx.x * 8 + w) * rows_per_warp;
  // caller-supplied indices are read one iteration ahead: the index load (a DRAM round trip of its own) then overlaps the
  // row copies of the current iteration instead of preceding them (gather(idx) ran at 0.66 vs 0.78 for the Philox mode)
  int32_t r_pre = 0;
  if (idx && base + lane < B && lane < rows_per_warp) r_pre = __ldg(idx + base + lane);
  for (; base < B; base += step) {
    const int nrow = (int)min((int64_t)rows_per_warp, (int64_t)B - base);
    int64_t r = 0;
    if (lane < nrow) {
      const int b = (int)(base + lane);
      if (idx) r = r_pre;
      else {
        r = philox_index(seed, (uint32_t)counter, (uint32_t)b, (uint32_t)(counter >> 32) + 0x51u, (int)size);
        if (idx_out) idx_out[b] = (int32_t)r;
      }
    }
    if (idx && base + step + lane < B && lane < rows_per_warp) r_pre = __ldg(idx + base + step + lane);
    if (lane == 0) tc5::mbar_arrive_expect_tx(&bars[w], (uint32_t)nrow * row_bytes);
    __syncwarp();
    if (lane < nrow) tc5::bulk_g2s(sm_w + (uint32_t)lane * row_bytes, rows + r * stride, row_bytes, &bars[w]);
    float4 c = make_float4(0.f, 0.f, 0.f, 0.f);
    if (out_cold && lane < nrow) c = __ldg(reinterpret_cast<const float4*>(cold) + r);
    tc5::mbar_wait(&bars[w], phase);
    phase ^= 1u;
    if (lane == 0) {
      tc5::bulk_s2g(out_hot + (size_t)base * stride, sm_w, (uint32_t)nrow * row_bytes);
      tc5::bulk_commit();
      tc5::bulk_wait_read();          // the staging block may be refilled once the store has read it
    }
    if (out_cold && lane < nrow) reinterpret_cast<float4*>(out_cold)[base + lane] = c;
    __syncwarp();
  }
}

static void launch_gather(const ilsw_rb* rb, const int32_t* idx, int B, float* out_hot, float* out_cold, uint64_t seed,
                          uint64_t counter, int32_t* idx_out, cudaStream_t st) {
  static const bool legacy = getenv("ILSW_GATHER_LEGACY") != nullptr;   // development aid (tools/replay_bench.py): the register-path kernels
  // rows up to 512 B go through the TMA unit (one bulk copy per row: Hopper 128 B 0.78 vs 0.65 of the HBM roofline, Walker
  // 192 B 0.80 vs 0.55); longer rows (Ant 960 B, Humanoid 3136 B) keep a warp's 128-bit loads busy on their own and the
  // register path wins (0.90 vs 0.83, 0.93 vs 0.88) -- profiles/r2_replay_bench.txt
  if (!legacy && (rb->stride & 15) == 0 && rb->stride * 4 <= 512) {
    const int row_bytes = rb->stride * 4;
    int rpw = kGatherStageBytes / row_bytes;
    rpw = rpw < 1 ? 1 : (rpw > 32 ? 32 : rpw);
    const size_t smem = (size_t)8 * rpw * row_bytes;
    static int sms = 0;
    if (!sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev); }
    if (smem > 48 * 1024) cudaFuncSetAttribute(rb_gather_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int64_t warps_needed = ((int64_t)B + rpw - 1) / rpw;
    int64_t grid = (warps_needed + 7) / 8;
    const int64_t cap = (int64_t)sms * 6;      // persistent-ish: up to 6 CTAs (48 warps, <= 192 KB of staged rows) per SM
    if (grid > cap) grid = cap;
    rb_gather_bulk_kernel<<<(unsigned)grid, 256, smem, st>>>(rb->rows, rb->cold, idx, B, rb->stride, out_hot, out_cold, rb->size, seed, counter, idx_out, rpw);
    return;
  }
  const int nv = rb->stride >> 2;           // 16-byte vectors per row: Hopper 7, Walker 11, Ant 58, Humanoid 193
  static const bool v1 = getenv("ILSW_GATHER_V1") != nullptr;   // development aid (tools/replay_bench.py): warp per row, no batching
  if (v1)
    rb_gather_kernel<32, 1><<<(B + 7) / 8, 256, 0, st>>>(rb->rows, rb->cold, idx, B, rb->stride, out_hot, out_cold, rb->size, seed, counter, idx_out);
  else if (nv <= 8)
    rb_gather_kernel<8, 1><<<(B + 31) / 32, 256, 0, st>>>(rb->rows, rb->cold, idx, B, rb->stride, out_hot, out_cold, rb->size, seed, counter, idx_out);
  else if (nv <= 16)
    rb_gather_kernel<16, 1><<<(B + 15) / 16, 256, 0, st>>>(rb->rows, rb->cold, idx, B, rb->stride, out_hot, out_cold, rb->size, seed, counter, idx_out);
  else {
    static const int un = getenv("ILSW_GATHER_UNROLL") ? atoi(getenv("ILSW_GATHER_UNROLL")) : 0;
    const int u = un ? un : (nv <= 64 ? 2 : 1);   // measured (profiles/r1_replay_bench.txt): Ant rows (58 vectors) 2 loads in flight per lane, Humanoid 1
    if (u == 1) rb_gather_kernel<32, 1><<<(B + 7) / 8, 256, 0, st>>>(rb->rows, rb->cold, idx, B, rb->stride, out_hot, out_cold, rb->size, seed, counter, idx_out);
    else if (u == 2) rb_gather_kernel<32, 2><<<(B + 7) / 8, 256, 0, st>>>(rb->rows, rb->cold, idx, B, rb->stride, out_hot, out_cold, rb->size, seed, counter, idx_out);
    else if (u == 4) rb_gather_kernel<32, 4><<<(B + 7) / 8, 256, 0, st>>>(rb->rows, rb->cold, idx, B, rb->stride, out_hot, out_cold, rb->size, seed, counter, idx_out);
    else rb_gather_kernel<32, 8><<<(B + 7) / 8, 256, 0, st>>>(rb->rows, rb->cold, idx, B, rb->stride, out_hot, out_cold, rb->size, seed, counter, idx_out);
  }
}

extern "C" int ilsw_rb_gather(ilsw_rb* rb, const int32_t* idx_dev, int B, float* out_hot, float* out_cold, void* stream) {
  if (!rb || !idx_dev || !out_hot || B <= 0) return fail(ILSW_ERR_ARG, "rb_gather: bad arguments");
  int rc = ilsw_rb_commit(rb, stream);
  if (rc) return rc;
  if (rb->size == 0) return fail(ILSW_ERR_STATE, "rb_gather: empty ring");
  launch_gather(rb, idx_dev, B, out_hot, out_cold, 0, 0, nullptr, (cudaStream_t)stream);
  CU(cudaGetLastError());
  return ILSW_OK;
}

extern "C" int ilsw_rb_sample(ilsw_rb* rb, int B, uint64_t seed, uint64_t counter, int32_t* idx_out_dev, float* out_hot,
                              void* stream) {
  if (!rb || !out_hot || B <= 0) return fail(ILSW_ERR_ARG, "rb_sample: bad arguments");
  int rc = ilsw_rb_commit(rb, stream);
  if (rc) return rc;
  if (rb->size == 0) return fail(ILSW_ERR_STATE, "rb_sample: empty ring");
  launch_gather(rb, nullptr, B, out_hot, nullptr, seed, counter, idx_out_dev, (cudaStream_t)stream);
  CU(cudaGetLastError());
  return ILSW_OK;
}

// ==========================================================================================
// Trainer
// ==========================================================================================
struct ilsw_trainer {
  TrainerSpec spec;
  Program host_prog;      // device pointers inside
  Program* dev_prog;
  char* scratch;
  size_t scratch_bytes;
  BarrierState* bar;      // [2]: launches alternate (bar_cur), each zeroes the other one
  int bar_cur;
  int bar_dirty;          // an aborted launch may leave its barrier state non-zero: the next launch memsets both
  void* mail_host;        // pinned mapped mailbox: [0] u64 done sequence, [64..] loss rows of the latest launch
  void* mail_dev;
  unsigned long long mail_seq;     // sequence number of the latest launch
  int mail_steps;                  // its step count
  int grid;
  int ctas;
  int tc5;                // engine variant with the tcgen05/TMA GEMM tile (ilsw_tc5.cuh)
  int sms;
  void* tmaps;            // device CUtensorMap[2 * kMaxOps]: A / B operand maps of the tcgen05 GEMM ops
  size_t smem_bytes;
  int t[kMaxNets];
  int n_total;
  int last_steps;
  int64_t launches;
  int profile;
  // replicas
  char* ipc_buf;          // [flags 256B][recv 2*8*n floats]
  size_t ipc_bytes;
  Replica rep;
  unsigned seq;
  void* peer_bases[8];
  // sampler-side inference with host buffers (ilsw_policy_act_host): pinned + device staging, allocated on first use
  float *act_pin, *act_dev;     // [kMaxActRows x (in_dim + act_dim)] each: observations first, actions after
  int update_mode;              // UpdateMode of the next launches (AdvIRL programs)
  HerSampling her;              // relabel-at-sample of the next launches (ilsw_trainer_set_her)
};
static const int kMaxActRows = 4096;

static const void* engine_fn(const ilsw_trainer* tr) {
  if (tr->tc5) return (const void*)ilsw_engine_kernel<1, true>;
  return tr->ctas == 2 ? (const void*)ilsw_engine_kernel<2, false> : (const void*)ilsw_engine_kernel<1, false>;
}
static size_t engine_smem(const ilsw_trainer* tr, int n_ops) {
  return engine_staging_bytes(tr->ctas, tr->tc5 != 0) + ((sizeof(Phase) * (size_t)tr->host_prog.n_phases + 15) & ~size_t(15)) +
         ((sizeof(Op) * (size_t)n_ops + 15) & ~size_t(15)) + sizeof(Ctx) + 64;
}
// occupancy variant + launch geometry of the program just built
static int engine_configure(ilsw_trainer* tr) {
  // 2 CTAs/SM pays off when the mma.sync GEMM phases have more than 2 tiles per SM (B >= 512); the tcgen05 variant is 1 CTA/SM
  tr->ctas = (!tr->tc5 && tr->spec.cfg.batch >= 512) ? 2 : 1;
  const char* cv = getenv("ILSW_CTAS_PER_SM");
  if (!tr->tc5 && cv && (atoi(cv) == 1 || atoi(cv) == 2)) tr->ctas = atoi(cv);
  tr->smem_bytes = engine_smem(tr, tr->host_prog.n_ops);     // the program is built: its op table is what the kernel copies
  const void* kfn = engine_fn(tr);
  cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tr->smem_bytes);
  if (e != cudaSuccess) return fail(ILSW_ERR_CUDA, "engine smem opt-in (%zu B): %s", tr->smem_bytes, cudaGetErrorString(e));
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, kThreads, tr->smem_bytes);
  if (e != cudaSuccess || per_sm < 1) return fail(ILSW_ERR_CUDA, "engine kernel cannot be resident (%s)", cudaGetErrorString(e));
  if (per_sm < tr->ctas) return fail(ILSW_ERR_CUDA, "engine kernel: only %d CTA/SM resident, need %d", per_sm, tr->ctas);
  tr->grid = tr->sms * tr->ctas;  // persistent: all CTAs co-resident (cooperative launch)
  const char* g = getenv("ILSW_GRID");
  if (g && atoi(g) > 0 && atoi(g) <= tr->sms * tr->ctas) tr->grid = atoi(g);
  return ILSW_OK;
}

static int trainer_build(ilsw_trainer* tr) {
  std::string why;
  int rc = validate_spec(tr->spec, &why);
  if (rc) return fail(rc, "trainer spec invalid: %s", why.c_str());
  const char* tv = getenv("ILSW_TC5");
  tr->tc5 = (tc5_wanted(tr->spec) && tmap_encoder() != nullptr && !(tv && atoi(tv) == 0)) ? 1 : 0;
  Bump measure;
  Program* tmp = new Program();
  const char* fv = getenv("ILSW_FUSE_L0");
  // first-layer fusion is OFF by default: measured on the B200 it lengthens the fused phases more than the saved phase +
  // barrier gives back (SAC Hopper 149.9 vs 147.0 us/step, GAIL Walker 272.8 vs 263.9; profiles/r2_ab_fuse_l0.txt)
  const bool fuse_l0 = fv ? atoi(fv) != 0 : false;
  rc = assemble(*tmp, tr->spec, measure, tr->tc5 != 0, fuse_l0);
  delete tmp;
  if (rc) return fail(rc, "program assembly failed (too many phases/ops?)");
  // a rebuild (attach_disc after load_snapshot) keeps the device-resident optimiser state of alpha
  DynState old_dyn;
  bool have_old = false;
  if (tr->scratch) {
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(&old_dyn, tr->host_prog.ctx.dyn, sizeof(old_dyn), cudaMemcpyDeviceToHost));
    have_old = true;
    CU(cudaFree(tr->scratch)); tr->scratch = nullptr;
  }
  tr->scratch_bytes = measure.off + 256;
  CU(cudaMalloc(&tr->scratch, tr->scratch_bytes));
  CU(cudaMemset(tr->scratch, 0, tr->scratch_bytes));
  Bump mem;
  mem.base = tr->scratch;
  rc = assemble(tr->host_prog, tr->spec, mem, tr->tc5 != 0, fuse_l0);
  if (rc) return fail(rc, "program assembly failed");
  // TMA panels of the mma.sync tile (tensor-core precision modes): every operand TMA can address gets a map with
  // {32 floats, 32 rows} boxes (k-contiguous: plain 128-byte swizzle; m/n-contiguous: 32-byte atoms, see gemm_tile_tc)
  const char* tpv = getenv("ILSW_TMA_PANELS");
  const bool tma_panels = tmap_encoder() != nullptr && tr->spec.cfg.gemm_precision != 0 && !(tpv && atoi(tpv) == 0);
  if (tr->tc5 || tma_panels) {           // TMA tensor maps of the GEMM operands (tcgen05 box shapes: ilsw_tc5.cuh)
    static_assert(sizeof(CUtensorMap) == 128, "tensor map size");
    if (!tr->tmaps) CU(cudaMalloc(&tr->tmaps, sizeof(CUtensorMap) * 2 * kMaxOps));
    std::vector<CUtensorMap> maps(2 * kMaxOps);
    Program& P = tr->host_prog;
    for (int i = 0; i < P.n_ops; ++i) {
      if (P.ops[i].kind != OP_GEMM) continue;
      if (!P.ops[i].gemm.tc5) {
        GemmOp& g = P.ops[i].gemm;
        g.tma = 0;
        if (!tma_panels || gemm_is_skinny(g)) continue;
        if (!g.a0 && tmap_operand_ok(g.A, g.lda) &&
            (g.a_mc ? make_tmap_2d(&maps[2 * i], g.A, g.lda, g.M, g.K, 32, true) : make_tmap_2d(&maps[2 * i], g.A, g.lda, g.K, g.M, 32, false)) == 0)
          g.tma |= 1;
        if (tmap_operand_ok(g.B, g.ldb) &&
            (g.b_nc ? make_tmap_2d(&maps[2 * i + 1], g.B, g.ldb, g.N, g.K, 32, true) : make_tmap_2d(&maps[2 * i + 1], g.B, g.ldb, g.K, g.N, 32, false)) == 0)
          g.tma |= 2;
        g.tmapA = reinterpret_cast<const CUtensorMap*>(tr->tmaps) + 2 * i;
        g.tmapB = reinterpret_cast<const CUtensorMap*>(tr->tmaps) + 2 * i + 1;
        continue;
      }
      GemmOp& g = P.ops[i].gemm;
      const int ra = g.a_mc ? make_tmap_2d(&maps[2 * i], g.A, g.lda, g.M, g.K, 32, true) : make_tmap_2d(&maps[2 * i], g.A, g.lda, g.K, g.M, tc5::kBM, false);
      const int rb = g.b_nc ? make_tmap_2d(&maps[2 * i + 1], g.B, g.ldb, g.N, g.K, 32, true) : make_tmap_2d(&maps[2 * i + 1], g.B, g.ldb, g.K, g.N, kTc5BN, false);
      if (ra || rb) return fail(ILSW_ERR_CUDA, "tensor map encode failed for GEMM op %d (%dx%dx%d): %d %d", i, g.M, g.N, g.K, ra, rb);
      g.tmapA = reinterpret_cast<const CUtensorMap*>(tr->tmaps) + 2 * i;
      g.tmapB = reinterpret_cast<const CUtensorMap*>(tr->tmaps) + 2 * i + 1;
    }
    CU(cudaMemcpy(tr->tmaps, maps.data(), sizeof(CUtensorMap) * 2 * kMaxOps, cudaMemcpyHostToDevice));
  }
  // development aid (tools/phase_profile.py): ILSW_DEBUG_DUP_PHASE=k runs phase k twice in a row -- only meaningful for
  // idempotent phases (forward / row phases); the second run shows the phase's warm-code, warm-data time
  if (const char* dp = getenv("ILSW_DEBUG_DUP_PHASE")) {
    Program& P = tr->host_prog;
    const int k = atoi(dp);
    if (k >= 0 && k < P.n_phases && P.n_phases < kMaxPhases) {
      for (int i = P.n_phases; i > k; --i) P.phases[i] = P.phases[i - 1];
      P.n_phases++;
    }
  }
  if (!tr->dev_prog) CU(cudaMalloc(&tr->dev_prog, sizeof(Program)));
  CU(cudaMemcpy(tr->dev_prog, &tr->host_prog, sizeof(Program), cudaMemcpyHostToDevice));
  DynState d;
  memset(&d, 0, sizeof(d));
  d.log_alpha = log(tr->spec.cfg.alpha > 0 ? tr->spec.cfg.alpha : 1.0);
  d.alpha = (float)exp(d.log_alpha);
  d.alpha_p1 = d.alpha_p2 = 1.0;
  if (have_old) { d = old_dyn; d.abort_flag = 0; d.error_code = 0; }
  CU(cudaMemcpy(tr->host_prog.ctx.dyn, &d, sizeof(d), cudaMemcpyHostToDevice));
  return engine_configure(tr);
}

extern "C" int ilsw_trainer_create(ilsw_trainer** out, const ilsw_trainer_config* cfg, const ilsw_mlp* nets, int n_nets) {
  if (!out || !cfg || !nets || n_nets <= 0 || n_nets > 6) return fail(ILSW_ERR_ARG, "trainer_create: bad arguments");
  int dev = 0, coop = 0, sms = 0;
  CU(cudaGetDevice(&dev));
  CU(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  if (!coop) return fail(ILSW_ERR_UNSUPPORTED, "device lacks cooperative launch");
  ilsw_trainer* tr = new ilsw_trainer();
  memset(tr, 0, sizeof(*tr));
  tr->spec.cfg = *cfg;
  for (int i = 0; i < n_nets; ++i) tr->spec.nets[i] = nets[i];
  tr->spec.n_nets = n_nets;
  tr->sms = sms;
  int rc = trainer_build(tr);
  if (rc) { ilsw_trainer_destroy(tr); return rc; }
  // two barrier states used by alternate launches: a launch zeroes the one its successor will use (no memset node per launch)
  cudaError_t e = cudaMalloc(&tr->bar, 2 * sizeof(BarrierState));
  if (e == cudaSuccess) e = cudaMemset(tr->bar, 0, 2 * sizeof(BarrierState));
  // host mailbox for the loss log (pinned + mapped; see RunArgs::mail_losses)
  if (e == cudaSuccess) {
    const size_t mb = 64 + (size_t)cfg->max_steps_per_call * kLossSlots * sizeof(float);
    e = cudaHostAlloc(&tr->mail_host, mb, cudaHostAllocMapped | cudaHostAllocPortable);
    if (e == cudaSuccess) { memset(tr->mail_host, 0, mb); e = cudaHostGetDevicePointer(&tr->mail_dev, tr->mail_host, 0); }
    if (e != cudaSuccess) { tr->mail_host = nullptr; tr->mail_dev = nullptr; e = cudaSuccess; cudaGetLastError(); }   // optional: fall back to copies
  }
  if (e != cudaSuccess) { ilsw_trainer_destroy(tr); return fail(ILSW_ERR_CUDA, "barrier alloc: %s", cudaGetErrorString(e)); }
  tr->rep.world = 1;
  *out = tr;
  return ILSW_OK;
}

extern "C" int ilsw_trainer_attach_disc(ilsw_trainer* tr, const ilsw_disc_config* cfg, const ilsw_mlp* disc) {
  if (!tr || !cfg || !disc) return fail(ILSW_ERR_ARG, "attach_disc: bad arguments");
  tr->spec.has_disc = 1;
  tr->spec.dcfg = *cfg;
  tr->spec.disc = *disc;
  return trainer_build(tr);
}

extern "C" int ilsw_trainer_set_her(ilsw_trainer* tr, const ilsw_her_sampling* her) {
  if (!tr) return fail(ILSW_ERR_ARG, "set_her: null trainer");
  memset(&tr->her, 0, sizeof(tr->her));
  if (!her || !her->enabled) return ILSW_OK;
  if (tr->spec.has_disc) return fail(ILSW_ERR_UNSUPPORTED, "set_her: not with a discriminator attached");
  if (her->goal_dim <= 0 || her->goal_dim >= tr->spec.cfg.obs_dim || her->relabel_num < 0 || her->relabel_num > tr->spec.cfg.batch ||
      !her->traj_start || !her->traj_len || !her->next_achieved_goal || her->n_traj < 0)
    return fail(ILSW_ERR_ARG, "set_her: bad arguments");
  tr->her.enabled = 1; tr->her.n_traj = her->n_traj;
  tr->her.traj_start = her->traj_start; tr->her.traj_len = her->traj_len;
  tr->her.ag_next = her->next_achieved_goal; tr->her.G = her->goal_dim;
  tr->her.relabel_num = her->relabel_num; tr->her.threshold = her->distance_threshold;
  tr->her.inj_idx_her = her->inj_idx_her;
  return ILSW_OK;
}

extern "C" int ilsw_trainer_set_update_mode(ilsw_trainer* tr, int mode) {
  if (!tr || mode < UPDATE_BOTH || mode > UPDATE_POLICY_ONLY) return fail(ILSW_ERR_ARG, "set_update_mode: bad arguments");
  if (mode != UPDATE_BOTH && !tr->spec.has_disc) return fail(ILSW_ERR_STATE, "set_update_mode: no discriminator attached");
  if (mode != UPDATE_BOTH && tr->rep.world > 1) return fail(ILSW_ERR_UNSUPPORTED, "set_update_mode: split launches are single-replica");
  tr->update_mode = mode;
  return ILSW_OK;
}

extern "C" int ilsw_trainer_destroy(ilsw_trainer* tr) {
  if (!tr) return ILSW_OK;
  cudaDeviceSynchronize();
  for (int r = 0; r < 8; ++r)
    if (tr->peer_bases[r] && r != tr->rep.rank) cudaIpcCloseMemHandle(tr->peer_bases[r]);
  if (tr->ipc_buf) cudaFree(tr->ipc_buf);
  if (tr->act_pin) cudaFreeHost(tr->act_pin);
  if (tr->act_dev) cudaFree(tr->act_dev);
  if (tr->scratch) cudaFree(tr->scratch);
  if (tr->tmaps) cudaFree(tr->tmaps);
  if (tr->dev_prog) cudaFree(tr->dev_prog);
  if (tr->bar) cudaFree(tr->bar);
  if (tr->mail_host) cudaFreeHost(tr->mail_host);
  delete tr;
  return ILSW_OK;
}

extern "C" int ilsw_train(ilsw_trainer* tr, ilsw_rb* policy_rb, ilsw_rb* expert_rb, int n_steps, const ilsw_inject* inject,
                          const ilsw_batch* batch, uint64_t seed, int stats_step, void* stream) {
  if (!tr || n_steps <= 0) return fail(ILSW_ERR_ARG, "train: bad arguments");
  const ilsw_trainer_config& cfg = tr->spec.cfg;
  if (n_steps > cfg.max_steps_per_call) return fail(ILSW_ERR_ARG, "train: n_steps %d > max_steps_per_call %d", n_steps, cfg.max_steps_per_call);
  if (batch && n_steps != 1) return fail(ILSW_ERR_ARG, "train: a caller-provided batch implies n_steps == 1");
  if (batch && tr->spec.has_disc) return fail(ILSW_ERR_UNSUPPORTED, "train: direct batches are not supported with a discriminator");
  if (!batch && !policy_rb) return fail(ILSW_ERR_ARG, "train: no replay ring and no batch");
  if (tr->spec.has_disc && !expert_rb) return fail(ILSW_ERR_ARG, "train: AdvIRL engine needs the expert ring");
  cudaStream_t st = (cudaStream_t)stream;
  RunArgs a;
  memset(&a, 0, sizeof(a));
  if (policy_rb) {
    if (policy_rb->O != cfg.obs_dim || policy_rb->A != cfg.act_dim) return fail(ILSW_ERR_ARG, "train: ring dims != trainer dims");
    int rc = ilsw_rb_commit(policy_rb, stream);
    if (rc) return rc;
    if (!batch && policy_rb->size <= 0) return fail(ILSW_ERR_STATE, "train: replay ring is empty");
    a.ring_policy.rows = policy_rb->rows; a.ring_policy.stride = policy_rb->stride; a.ring_policy.size = (int)policy_rb->size;
  }
  if (expert_rb) {
    if (expert_rb->O != cfg.obs_dim || expert_rb->A != cfg.act_dim) return fail(ILSW_ERR_ARG, "train: expert ring dims != trainer dims");
    int rc = ilsw_rb_commit(expert_rb, stream);
    if (rc) return rc;
    if (tr->spec.has_disc && expert_rb->size <= 0) return fail(ILSW_ERR_STATE, "train: expert ring is empty");
    a.ring_expert.rows = expert_rb->rows; a.ring_expert.stride = expert_rb->stride; a.ring_expert.size = (int)expert_rb->size;
  }
  a.n_steps = n_steps; a.step0 = tr->n_total; a.stats_step = stats_step; a.seed = seed;
  for (int i = 0; i < kMaxNets; ++i) a.t0[i] = tr->t[i];
  if (inject) {
    const bool policy_part = tr->update_mode != UPDATE_DISC_ONLY;
    if (policy_part && (!inject->idx || !inject->eps_next)) return fail(ILSW_ERR_ARG, "train: inject needs idx and eps_next");
    if (policy_part && cfg.algo == ILSW_ALGO_SAC_ALPHA && !inject->eps_cur) return fail(ILSW_ERR_ARG, "train: inject needs eps_cur for SAC");
    if (tr->spec.has_disc && tr->update_mode != UPDATE_POLICY_ONLY &&
        (!inject->idx_expert || !inject->idx_policy_d || (tr->spec.dcfg.use_grad_pen && !inject->gp_eps)))
      return fail(ILSW_ERR_ARG, "train: inject needs idx_expert/idx_policy_d/gp_eps for the discriminator step");
    a.has_inject = 1;
    a.inj.idx = inject->idx; a.inj.eps_next = inject->eps_next; a.inj.eps_cur = inject->eps_cur;
    a.inj.idx_expert = inject->idx_expert; a.inj.idx_policy_d = inject->idx_policy_d; a.inj.gp_eps = inject->gp_eps;
  }
  if (batch) {
    if (!batch->obs || !batch->act || !batch->rew || !batch->term || !batch->next_obs) return fail(ILSW_ERR_ARG, "train: incomplete batch");
    a.has_direct = 1;
    a.direct.obs = batch->obs; a.direct.act = batch->act; a.direct.rew = batch->rew; a.direct.term = batch->term;
    a.direct.next_obs = batch->next_obs;
    if (!inject) a.has_inject = 0;
  }
  a.world = tr->rep.world; a.rank = tr->rep.rank; a.loss_log_offset = 0;
  a.profile = tr->profile;
  a.update_mode = tr->spec.has_disc ? tr->update_mode : UPDATE_BOTH;
  a.her = tr->her;
  if (a.her.enabled && !batch) {
    if (!policy_rb) return fail(ILSW_ERR_ARG, "train: hindsight sampling needs the replay ring");
    if (a.her.n_traj <= 0) return fail(ILSW_ERR_STATE, "train: hindsight sampling needs at least one finished trajectory");
    if (inject && !a.her.inj_idx_her) return fail(ILSW_ERR_ARG, "train: inject needs inj_idx_her for hindsight sampling");
  }
  if (a.update_mode == UPDATE_DISC_ONLY && batch) return fail(ILSW_ERR_ARG, "train: a disc-only launch samples from the rings (no direct batch)");
  Replica rp = tr->rep;
  rp.seq0 = tr->seq;
  rp.grad = tr->host_prog.ctx.policy.g;
  rp.g_splits = tr->host_prog.ctx.policy.g_splits; rp.g_split_stride = tr->host_prog.ctx.policy.g_stride;
  const Program* dp = tr->dev_prog;
  if (tr->bar_dirty) { CU(cudaMemsetAsync(tr->bar, 0, 2 * sizeof(BarrierState), st)); tr->bar_dirty = 0; }
  BarrierState* bar = tr->bar + tr->bar_cur;               // zeroed by the previous launch (or at creation)
  a.bar_other = &tr->bar[tr->bar_cur ^ 1].count;
  tr->bar_cur ^= 1;
  if (tr->mail_dev) {
    tr->mail_seq += 1; tr->mail_steps = n_steps;
    a.mail_done = reinterpret_cast<unsigned long long*>(tr->mail_dev);
    a.mail_losses = reinterpret_cast<float*>(reinterpret_cast<char*>(tr->mail_dev) + 64);
    a.mail_seq = tr->mail_seq;
  }
  void* args[] = {(void*)&dp, (void*)&a, (void*)&bar, (void*)&rp};
  const size_t smem = engine_smem(tr, tr->host_prog.n_ops);
  CU(cudaLaunchCooperativeKernel(engine_fn(tr), dim3(tr->grid), dim3(kThreads), args, smem, st));
  tr->launches += 1;
  // host mirrors of the on-device counters
  if (a.update_mode != UPDATE_DISC_ONLY)
    tr->seq += (unsigned)(adam_t(a, tr->host_prog.ctx.hp, SLOT_POLICY, n_steps - 1) - a.t0[SLOT_POLICY]);
  commit_counters(tr->t, tr->n_total, a, tr->host_prog.ctx.hp, n_steps);
  tr->last_steps = n_steps;
  return ILSW_OK;
}

static int check_abort(ilsw_trainer* tr, cudaStream_t st) {
  DynState d;
  CU(cudaMemcpyAsync(&d, tr->host_prog.ctx.dyn, sizeof(d), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  if (d.abort_flag) { tr->bar_dirty = 1; return fail(ILSW_ERR_ABORTED, "engine launch aborted (barrier/replica wait timed out)"); }
  return ILSW_OK;
}
// losses of the LATEST launch through the host mailbox: poll the sequence word the kernel publishes after its last step.
// Returns 1 when served, 0 when the caller must take the copy path (no mailbox, other step count, or ~2 s without progress).
static int mailbox_losses(ilsw_trainer* tr, float* host_out, int n_steps) {
  if (!tr->mail_host || tr->mail_seq == 0 || n_steps > tr->mail_steps) return 0;
  volatile unsigned long long* done = reinterpret_cast<volatile unsigned long long*>(tr->mail_host);
  const unsigned long long want = tr->mail_seq;
  const auto t0 = std::chrono::steady_clock::now();
  for (long long spin = 0; *done < want; ++spin) {
    if ((spin & 4095) == 4095 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(6)) return 0;   // the engine's own watchdog fires after ~4 s
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
  __atomic_thread_fence(__ATOMIC_ACQUIRE);
  memcpy(host_out, reinterpret_cast<char*>(tr->mail_host) + 64, (size_t)n_steps * kLossSlots * sizeof(float));
  return 1;
}

extern "C" int ilsw_read_losses(ilsw_trainer* tr, float* host_out, int n_steps, void* stream) {
  if (!tr || !host_out || n_steps <= 0 || n_steps > tr->spec.cfg.max_steps_per_call) return fail(ILSW_ERR_ARG, "read_losses: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  if (mailbox_losses(tr, host_out, n_steps)) return ILSW_OK;       // a launch that aborted never publishes: copy path below reports it
  CU(cudaMemcpyAsync(host_out, tr->host_prog.ctx.loss_log, (size_t)n_steps * kLossSlots * sizeof(float), cudaMemcpyDeviceToHost, st));
  return check_abort(tr, st);
}
extern "C" int ilsw_read_losses_async(ilsw_trainer* tr, float* pinned_out, int n_steps, void* stream) {
  if (!tr || !pinned_out || n_steps <= 0 || n_steps > tr->spec.cfg.max_steps_per_call) return fail(ILSW_ERR_ARG, "read_losses_async: bad arguments");
  CU(cudaMemcpyAsync(pinned_out, tr->host_prog.ctx.loss_log, (size_t)n_steps * kLossSlots * sizeof(float), cudaMemcpyDeviceToHost,
                     (cudaStream_t)stream));
  return ILSW_OK;
}
extern "C" int ilsw_check_abort(ilsw_trainer* tr, void* stream) {
  if (!tr) return fail(ILSW_ERR_ARG, "check_abort: null");
  return check_abort(tr, (cudaStream_t)stream);
}
extern "C" int ilsw_stats_floats(const ilsw_trainer* tr) { return tr ? tr->host_prog.ctx.stats_floats : ILSW_ERR_ARG; }
extern "C" int ilsw_read_stats(ilsw_trainer* tr, float* host_out, int n_floats, void* stream) {
  if (!tr || !host_out || n_floats <= 0 || n_floats > tr->host_prog.ctx.stats_floats) return fail(ILSW_ERR_ARG, "read_stats: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaMemcpyAsync(host_out, tr->host_prog.ctx.stats, (size_t)n_floats * sizeof(float), cudaMemcpyDeviceToHost, st));
  return check_abort(tr, st);
}

extern "C" int ilsw_get_state(ilsw_trainer* tr, ilsw_state* out, void* stream) {
  if (!tr || !out) return fail(ILSW_ERR_ARG, "get_state: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  DynState d;
  CU(cudaMemcpyAsync(&d, tr->host_prog.ctx.dyn, sizeof(d), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  memset(out, 0, sizeof(*out));
  out->log_alpha = d.log_alpha; out->alpha_exp_avg = d.alpha_m; out->alpha_exp_avg_sq = d.alpha_v; out->alpha_step = d.alpha_t;
  for (int i = 0; i < 8; ++i) out->adam_step[i] = tr->t[i];
  out->n_train_steps_total = tr->n_total;
  if (d.abort_flag) return fail(ILSW_ERR_ABORTED, "engine launch aborted");
  return ILSW_OK;
}
extern "C" int ilsw_set_state(ilsw_trainer* tr, const ilsw_state* in, void* stream) {
  if (!tr || !in) return fail(ILSW_ERR_ARG, "set_state: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  DynState d;
  memset(&d, 0, sizeof(d));
  d.log_alpha = in->log_alpha; d.alpha_m = in->alpha_exp_avg; d.alpha_v = in->alpha_exp_avg_sq; d.alpha_t = in->alpha_step;
  d.alpha = (float)exp(d.log_alpha);
  d.alpha_p1 = pow(tr->host_prog.ctx.hp.beta1, (double)d.alpha_t);
  d.alpha_p2 = pow(tr->host_prog.ctx.hp.beta2, (double)d.alpha_t);
  CU(cudaMemcpyAsync(tr->host_prog.ctx.dyn, &d, sizeof(d), cudaMemcpyHostToDevice, st));
  CU(cudaStreamSynchronize(st));
  for (int i = 0; i < 8; ++i) tr->t[i] = in->adam_step[i];
  tr->n_total = in->n_train_steps_total;
  return ILSW_OK;
}

extern "C" int ilsw_describe_program(const ilsw_trainer* tr, char* buf, int buf_len) {
  if (!tr || !buf || buf_len <= 0) return fail(ILSW_ERR_ARG, "describe_program: bad arguments");
  std::string s = describe_program(tr->host_prog);
  snprintf(buf, buf_len, "%s", s.c_str());
  return (int)s.size();
}
extern "C" int ilsw_read_phase_ns(ilsw_trainer* tr, unsigned long long* host_out, int n, void* stream) {
  if (!tr || !host_out || n <= 0 || n > 2 * (kMaxPhases + 1)) return fail(ILSW_ERR_ARG, "read_phase_ns: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaMemcpyAsync(host_out, tr->host_prog.ctx.phase_ns, (size_t)n * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return ILSW_OK;
}
extern "C" int ilsw_read_cta_ns(ilsw_trainer* tr, unsigned long long* host_out /* [96][304] */, void* stream) {
  if (!tr || !host_out) return fail(ILSW_ERR_ARG, "read_cta_ns: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  CU(cudaMemcpyAsync(host_out, tr->host_prog.ctx.cta_ns, sizeof(unsigned long long) * kMaxPhases * kMaxGrid, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  return ILSW_OK;
}
extern "C" int ilsw_read_tile_ns(unsigned long long* host_out /* [kMaxPhases][8] */) {
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpyFromSymbol(host_out, g_tile_ns, sizeof(unsigned long long) * kMaxPhases * 8));
  return ILSW_OK;
}
extern "C" int ilsw_trainer_set_profiling(ilsw_trainer* tr, int on) {
  if (!tr) return fail(ILSW_ERR_ARG, "set_profiling: null");
  tr->profile = on ? 1 : 0;
  return ILSW_OK;
}
extern "C" int ilsw_num_phases(const ilsw_trainer* tr) { return tr ? tr->host_prog.n_phases : ILSW_ERR_ARG; }
extern "C" int64_t ilsw_kernel_launches(const ilsw_trainer* tr) { return tr ? tr->launches : ILSW_ERR_ARG; }
extern "C" int ilsw_trainer_uses_tc5(const ilsw_trainer* tr) { return tr ? tr->tc5 : ILSW_ERR_ARG; }

extern "C" int ilsw_policy_act(ilsw_trainer* tr, const float* obs_dev, int n, int deterministic, uint64_t seed, float* act_dev,
                               void* stream) {
  if (!tr || !obs_dev || !act_dev || n <= 0 || n > 4096) return fail(ILSW_ERR_ARG, "policy_act: bad arguments");
  const MlpPtrs& P = tr->host_prog.ctx.policy;
  const Hyper& hp = tr->host_prog.ctx.hp;
  size_t sh = (size_t)(P.in_dim + 2 * P.hid) * sizeof(float);
  ilsw_policy_act_kernel<<<n, 256, sh, (cudaStream_t)stream>>>(P, hp.algo, hp.max_act, hp.policy_noise, hp.noise_clip, obs_dev, n,
                                                               deterministic, seed, act_dev);
  CU(cudaGetLastError());
  return ILSW_OK;
}

// A1 with HOST buffers -- what exploration_policy.get_actions(obs_np) does every env step (rlkit/torch/core.py:74-89,
// base_algorithm.py:369-380): pinned H2D of the observations, ONE kernel, pinned D2H of the actions, ONE host sync.
// Stream ordered behind the gradient steps already queued on `stream`, so it always sees the latest policy.
extern "C" int ilsw_policy_act_host(ilsw_trainer* tr, const float* obs_host, int n, int deterministic, uint64_t seed,
                                    float* act_host, void* stream) {
  if (!tr || !obs_host || !act_host || n <= 0 || n > kMaxActRows) return fail(ILSW_ERR_ARG, "policy_act_host: bad arguments");
  const MlpPtrs& P = tr->host_prog.ctx.policy;
  const int O = P.in_dim, A = tr->spec.cfg.act_dim;
  if (!tr->act_pin) {
    const size_t bytes = (size_t)kMaxActRows * (O + A) * sizeof(float);
    CU(cudaMallocHost(&tr->act_pin, bytes));
    CU(cudaMalloc(&tr->act_dev, bytes));
  }
  cudaStream_t st = (cudaStream_t)stream;
  float* pin_obs = tr->act_pin; float* pin_act = tr->act_pin + (size_t)kMaxActRows * O;
  float* dev_obs = tr->act_dev; float* dev_act = tr->act_dev + (size_t)kMaxActRows * O;
  memcpy(pin_obs, obs_host, (size_t)n * O * sizeof(float));
  CU(cudaMemcpyAsync(dev_obs, pin_obs, (size_t)n * O * sizeof(float), cudaMemcpyHostToDevice, st));
  int rc = ilsw_policy_act(tr, dev_obs, n, deterministic, seed, dev_act, stream);
  if (rc) return rc;
  CU(cudaMemcpyAsync(pin_act, dev_act, (size_t)n * A * sizeof(float), cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  memcpy(act_host, pin_act, (size_t)n * A * sizeof(float));
  return ILSW_OK;
}

// ==========================================================================================
// Replicas: CUDA-IPC mapped receive buffers + flags, exchanged inside the engine kernel
// ==========================================================================================
static const size_t kFlagBytes = 256;

extern "C" int ilsw_replica_export(ilsw_trainer* tr, void* handle_out) {
  if (!tr || !handle_out) return fail(ILSW_ERR_ARG, "replica_export: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) <= ILSW_IPC_HANDLE_BYTES, "handle size");
  const int n = tr->host_prog.ctx.policy.n_params;
  if (!tr->ipc_buf) {
    tr->ipc_bytes = kFlagBytes + (size_t)2 * 8 * round_up(n, 4) * sizeof(float);
    CU(cudaMalloc(&tr->ipc_buf, tr->ipc_bytes));
    CU(cudaMemset(tr->ipc_buf, 0, tr->ipc_bytes));
    CU(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, tr->ipc_buf));
  memset(handle_out, 0, ILSW_IPC_HANDLE_BYTES);
  memcpy(handle_out, &h, sizeof(h));
  return ILSW_OK;
}

// Exchange buffer provided by the caller (symmetric memory: the same allocation on every rank, mapped into every process,
// optionally with an NVLS multicast mapping).  Layout as in ilsw_replica_export: [kFlagBytes flags / counter][2][world][nstride].
extern "C" int64_t ilsw_replica_buffer_bytes(ilsw_trainer* tr) {
  if (!tr) return ILSW_ERR_ARG;
  return (int64_t)(kFlagBytes + (size_t)2 * 8 * round_up(tr->host_prog.ctx.policy.n_params, 4) * sizeof(float));
}
extern "C" int ilsw_replica_connect_symm(ilsw_trainer* tr, int rank, int world, const uint64_t* peer_ptrs, uint64_t multicast_ptr,
                                         int64_t bytes) {
  if (!tr || !peer_ptrs || world < 1 || world > 8 || rank < 0 || rank >= world) return fail(ILSW_ERR_ARG, "replica_connect_symm: bad arguments");
  if (bytes < ilsw_replica_buffer_bytes(tr)) return fail(ILSW_ERR_ARG, "replica_connect_symm: buffer of %lld B, need %lld", (long long)bytes, (long long)ilsw_replica_buffer_bytes(tr));
  const int n = tr->host_prog.ctx.policy.n_params;
  Replica& rp = tr->rep;
  memset(&rp, 0, sizeof(rp));
  rp.world = world; rp.rank = rank; rp.n = n; rp.nstride = round_up(n, 4);
  rp.grad = tr->host_prog.ctx.policy.g;
  for (int r = 0; r < world; ++r) {
    char* base = reinterpret_cast<char*>((uintptr_t)peer_ptrs[r]);
    if (!base || (reinterpret_cast<uintptr_t>(base) & 127)) return fail(ILSW_ERR_ARG, "replica_connect_symm: peer pointer %d null or unaligned", r);
    tr->peer_bases[r] = nullptr;                 // not ours to unmap
    rp.flags_peer[r] = reinterpret_cast<unsigned*>(base);
    rp.cnt_peer[r] = reinterpret_cast<unsigned long long*>(base + 128);
    rp.recv_peer[r] = reinterpret_cast<float*>(base + kFlagBytes);
  }
  rp.cnt_local = rp.cnt_peer[rank];
  rp.flags_local = rp.flags_peer[rank];
  rp.recv_local = rp.recv_peer[rank];
  rp.arrive_local = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(rp.flags_local) + 192);     // [192,196) of the flag area
  if (multicast_ptr) {
    char* mc = reinterpret_cast<char*>((uintptr_t)multicast_ptr);
    rp.cnt_mc = reinterpret_cast<unsigned long long*>(mc + 128);
    rp.recv_mc = reinterpret_cast<float*>(mc + kFlagBytes);
  }
  CU(cudaMemset(rp.flags_local, 0, kFlagBytes));
  CU(cudaDeviceSynchronize());
  tr->seq = 0;
  return ILSW_OK;
}

extern "C" int ilsw_replica_connect(ilsw_trainer* tr, int rank, int world, const void* all_handles) {
  if (!tr || !all_handles || world < 1 || world > 8 || rank < 0 || rank >= world) return fail(ILSW_ERR_ARG, "replica_connect: bad arguments");
  if (!tr->ipc_buf) return fail(ILSW_ERR_STATE, "replica_connect: call ilsw_replica_export first");
  const int n = tr->host_prog.ctx.policy.n_params;
  Replica& rp = tr->rep;
  memset(&rp, 0, sizeof(rp));
  CU(cudaMemset(tr->ipc_buf, 0, kFlagBytes));      // sequence flags and push counter restart with tr->seq (peers push only after the caller's barrier)
  CU(cudaDeviceSynchronize());
  rp.world = world; rp.rank = rank; rp.n = n; rp.nstride = round_up(n, 4);
  rp.grad = tr->host_prog.ctx.policy.g;
  for (int r = 0; r < world; ++r) {
    void* base = nullptr;
    if (r == rank) base = tr->ipc_buf;
    else {
      cudaIpcMemHandle_t h;
      memcpy(&h, (const char*)all_handles + (size_t)r * ILSW_IPC_HANDLE_BYTES, sizeof(h));
      CU(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    }
    tr->peer_bases[r] = base;
    rp.flags_peer[r] = reinterpret_cast<unsigned*>(base);
    rp.cnt_peer[r] = reinterpret_cast<unsigned long long*>((char*)base + 128);     // [0,32): sequence flags; 128: push counter
    rp.recv_peer[r] = reinterpret_cast<float*>((char*)base + kFlagBytes);
  }
  rp.cnt_local = rp.cnt_peer[rank];
  rp.flags_local = rp.flags_peer[rank];
  rp.recv_local = rp.recv_peer[rank];
  rp.arrive_local = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(rp.flags_local) + 192);
  tr->seq = 0;
  return ILSW_OK;
}

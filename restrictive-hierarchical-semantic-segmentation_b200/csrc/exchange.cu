// Data-parallel exchange buffer: the step's only cross-rank traffic is ONE all-reduce of
//   [ step summary (fp64, written by rhseg_step_finalize) | head / FiLM parameter gradients ]
// (rhseg_b200/dist.py).  These two kernels move the fp32 gradient tensors into / out of that fp64
// buffer in a single launch each, instead of one conversion per tensor plus a concatenation.
// The reference has no counterpart: it trains single-process (train.py:201-241); SURVEY.md 8(e).
#include <algorithm>
#include <cstring>
#include "common.cuh"

namespace rhseg {

constexpr int XCHG_MAX_PARTS = 32;

struct XchgParts {
  const float* src[XCHG_MAX_PARTS];
  float* dst[XCHG_MAX_PARTS];
  long end[XCHG_MAX_PARTS];  // exclusive prefix ends (elements)
  int n;
};

__device__ __forceinline__ int part_of(const XchgParts& p, long i) {
  int k = 0;
  while (k < p.n - 1 && i >= p.end[k]) ++k;
  return k;
}

__global__ void __launch_bounds__(256) pack_f64_kernel(XchgParts p, long total, double* __restrict__ out) {
  pdl_wait();
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
    const int k = part_of(p, i);
    const long base = k ? p.end[k - 1] : 0;
    out[i] = (double)__ldg(p.src[k] + (i - base));
  }
}

__global__ void __launch_bounds__(256) unpack_f32_kernel(XchgParts p, long total, const double* __restrict__ in, double scale) {
  pdl_wait();
  for (long i = (long)blockIdx.x * 256 + threadIdx.x; i < total; i += (long)gridDim.x * 256) {
    const int k = part_of(p, i);
    const long base = k ? p.end[k - 1] : 0;
    p.dst[k][i - base] = (float)(in[i] * scale);
  }
}

static int fill_parts(const void* const* ptrs, const long* counts, int n, bool is_dst, XchgParts& p, long& total) {
  if (!ptrs || !counts || n <= 0) return RHSEG_ERR_ARG;
  if (n > XCHG_MAX_PARTS) return RHSEG_ERR_UNSUPPORTED;
  total = 0;
  p.n = n;
  for (int k = 0; k < n; ++k) {
    if (counts[k] < 0 || (counts[k] > 0 && !ptrs[k])) return RHSEG_ERR_ARG;
    total += counts[k];
    p.end[k] = total;
    p.src[k] = is_dst ? nullptr : static_cast<const float*>(ptrs[k]);
    p.dst[k] = is_dst ? static_cast<float*>(const_cast<void*>(ptrs[k])) : nullptr;
  }
  return RHSEG_OK;
}

}  // namespace rhseg

using namespace rhseg;

extern "C" int rhseg_pack_f64(const void* const* srcs, const long* counts, int n, double* out, void* stream) {
  XchgParts p;
  long total = 0;
  const int rc = fill_parts(srcs, counts, n, false, p, total);
  if (rc != RHSEG_OK) return rc;
  if (!out) return RHSEG_ERR_ARG;
  if (total == 0) return RHSEG_OK;
  const unsigned grid = (unsigned)std::min<long>((total + 255) / 256, 4L * device_sm_count());
  launch_pdl(pack_f64_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, p, total, out);
  RHSEG_LAUNCH_CHECK();
  return RHSEG_OK;
}

extern "C" int rhseg_unpack_f32(const double* in, double scale, void* const* dsts, const long* counts, int n, void* stream) {
  XchgParts p;
  long total = 0;
  const int rc = fill_parts(dsts, counts, n, true, p, total);
  if (rc != RHSEG_OK) return rc;
  if (!in) return RHSEG_ERR_ARG;
  if (total == 0) return RHSEG_OK;
  const unsigned grid = (unsigned)std::min<long>((total + 255) / 256, 4L * device_sm_count());
  launch_pdl(unpack_f32_kernel, dim3(grid), dim3(256), 0, (cudaStream_t)stream, p, total, in, scale);
  RHSEG_LAUNCH_CHECK();
  return RHSEG_OK;
}

// ------------------------------------------------------------------------------------
// One-shot all-reduce over NVLink peer memory (single node, one process per GPU).
//
// The exchange buffer is ~100 KB: an NCCL all-reduce of that size is pure launch + protocol latency.
// Here every rank owns a receive area that all peers map through CUDA IPC:
//   recv[2 slots][world sources][cap] x 16 bytes   +   epoch[MAX_CTA] u32 | status u32
// ONE kernel per step converts the rank's contribution (summary + fp32 parameter gradients) to fp64
// and PUSHES every element into each peer's receive area as a self-validating 16-byte record
// { lo32, epoch, hi32, epoch } (each 8-byte half carries its own flag, so no fence, no separate flag
// and no ordering between stores is needed); it then polls its OWN memory until the peers' records
// of the same elements carry the current epoch and adds them in rank order (every rank adds in the
// same order: bit-identical results everywhere).  One NVLink one-way latency end to end; threads are
// independent (no CTA or grid synchronisation).  Two slots suffice: a peer can only be one epoch
// ahead, because finishing an epoch needs this rank's records of that epoch.  The epoch counters
// live in device memory, so the kernel replays inside a CUDA graph unchanged.
// ------------------------------------------------------------------------------------
namespace rhseg {

constexpr int XCHG_MAX_WORLD = 16, XCHG_MAX_CTA = 32, XCHG_THREADS = 256, XCHG_ELEMS = 2;
constexpr unsigned long long XCHG_DEFAULT_TIMEOUT_NS = 30ull * 1000 * 1000 * 1000;  // RHSEG_XCHG_TIMEOUT_MS / rhseg_xchg_set_timeout_ms

struct XchgDev {
  uint4* recv[XCHG_MAX_WORLD];  // peer r's receive area [2][world][cap]
  uint32_t* epoch;              // local [MAX_CTA]
  uint32_t* status;             // local, sticky: 0 ok, 1 = a wait for a peer timed out (every result from then on is NaN)
  unsigned long long timeout_ns;  // 0 = wait for ever (what an NCCL all-reduce does)
  long cap;
  int rank, world;
};

struct XchgCtx {
  XchgDev dev;
  void* local_base;
  void* peer_base[XCHG_MAX_WORLD];
  size_t recv_bytes;
  int world;
  bool connected;
  unsigned char own_handle[RHSEG_XCHG_HANDLE_BYTES];
};

__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

__device__ __forceinline__ void st_record(uint4* p, uint32_t lo, uint32_t hi, uint32_t ep) {
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(lo), "r"(ep), "r"(hi), "r"(ep) : "memory");
}
__device__ __forceinline__ uint4 ld_record(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(XCHG_THREADS)
xchg_all_reduce_kernel(XchgDev d, const double* summary, long n_sum, XchgParts p, long total, double* out) {
  pdl_wait();
  const int cta = blockIdx.x, tid = threadIdx.x;
  const uint32_t ep = d.epoch[cta] + 1u;  // every thread reads it before anybody of this CTA can have written it
  // A timeout is a hard failure: once a peer was missed the epochs of the ranks no longer agree, so this and every
  // later exchange of the context poisons its whole result with NaN (loss and gradients turn NaN visibly instead
  // of the replicas diverging silently); the host reads the sticky status at its next sync point (PeerExchange.check).
  bool failed = *reinterpret_cast<volatile uint32_t*>(d.status) != 0u;
  __syncthreads();
  const size_t slot_base = (size_t)(ep & 1u) * d.world * (size_t)d.cap;
  const long stride = (long)gridDim.x * XCHG_THREADS;
  for (long i0 = (long)cta * XCHG_THREADS + tid; i0 < total; i0 += stride * XCHG_ELEMS) {
    double v[XCHG_ELEMS];
    // 1. this rank's elements -> fp64 -> pushed into every peer's receive area
#pragma unroll
    for (int e = 0; e < XCHG_ELEMS; ++e) {
      const long i = i0 + e * stride;
      v[e] = 0.0;
      if (i < total) {
        if (i < n_sum) {
          v[e] = summary[i];
        } else {
          const long j = i - n_sum;
          const int k = part_of(p, j);
          v[e] = (double)__ldg(p.src[k] + (j - (k ? p.end[k - 1] : 0)));
        }
        const unsigned long long bits = (unsigned long long)__double_as_longlong(v[e]);
        for (int r = 0; r < d.world; ++r)
          if (r != d.rank) st_record(d.recv[r] + slot_base + (size_t)d.rank * d.cap + i, (uint32_t)bits, (uint32_t)(bits >> 32), ep);
      }
    }
    // 2. the peers' records of the same elements arrive in this rank's own memory; add in rank order
#pragma unroll
    for (int e = 0; e < XCHG_ELEMS; ++e) {
      const long i = i0 + e * stride;
      if (i < total) {
        double acc = 0.0;
        for (int r = 0; r < d.world; ++r) {
          double val = v[e];
          if (r != d.rank) {
            const uint4* src = d.recv[d.rank] + slot_base + (size_t)r * d.cap + i;
            uint4 rec = ld_record(src);
            if (!failed && (rec.y != ep || rec.w != ep)) {
              const unsigned long long t0 = d.timeout_ns ? global_ns() : 0ull;
              unsigned spins = 0;
              do {
                if (d.timeout_ns && (++spins & 1023u) == 0u && global_ns() - t0 > d.timeout_ns) {
                  *reinterpret_cast<volatile uint32_t*>(d.status) = 1u;
                  failed = true;
                  break;
                }
                rec = ld_record(src);
              } while (rec.y != ep || rec.w != ep);
            }
            val = __longlong_as_double((long long)(((unsigned long long)rec.z << 32) | rec.x));
          }
          acc += val;
        }
        out[i] = acc;
      }
    }
  }
  if (__syncthreads_or(failed ? 1 : 0)) {  // poison everything this CTA wrote (other CTAs see the sticky status themselves)
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    for (long i = (long)cta * XCHG_THREADS + tid; i < total; i += stride) out[i] = nan;
  }
  // all per-CTA counters advance together (CTA 0 also bumps those of the CTAs this launch did not use), so the
  // epoch is the number of exchanges done, whatever grid each of them ran with
  if (tid == 0) d.epoch[cta] = ep;
  if (cta == 0 && tid >= (int)gridDim.x && tid < XCHG_MAX_CTA) d.epoch[tid] = ep;
}

}  // namespace rhseg

extern "C" int rhseg_xchg_create(long capacity, int world, void** ctx_out, unsigned char* handle_out) {
  if (capacity <= 0 || world < 1 || !ctx_out || !handle_out) return RHSEG_ERR_ARG;
  if (world > XCHG_MAX_WORLD) return RHSEG_ERR_UNSUPPORTED;
  XchgCtx* c = new XchgCtx();
  c->world = world;
  c->recv_bytes = 2 * (size_t)world * (size_t)capacity * sizeof(uint4);
  const size_t bytes = c->recv_bytes + ((size_t)XCHG_MAX_CTA + 4) * sizeof(uint32_t);
  cudaError_t e = cudaMalloc(&c->local_base, bytes);
  if (e == cudaSuccess) e = cudaMemset(c->local_base, 0, bytes);  // epoch 0 everywhere: the first exchange uses epoch 1
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, c->local_base);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    if (c->local_base) cudaFree(c->local_base);
    delete c;
    return (int)e;
  }
  static_assert(sizeof(cudaIpcMemHandle_t) == RHSEG_XCHG_HANDLE_BYTES, "IPC handle size");
  memcpy(handle_out, &h, sizeof(h));
  memcpy(c->own_handle, &h, sizeof(h));
  c->dev.cap = capacity;
  c->dev.timeout_ns = XCHG_DEFAULT_TIMEOUT_NS;
  if (const char* e = getenv("RHSEG_XCHG_TIMEOUT_MS")) c->dev.timeout_ns = (unsigned long long)std::max(0L, atol(e)) * 1000000ull;
  c->connected = false;
  *ctx_out = c;
  return RHSEG_OK;
}

extern "C" int rhseg_xchg_connect(void* ctx, int rank, const unsigned char* handles) {
  XchgCtx* c = static_cast<XchgCtx*>(ctx);
  if (!c || !handles || rank < 0 || rank >= c->world) return RHSEG_ERR_ARG;
  if (c->connected) return RHSEG_ERR_ARG;
  for (int r = 0; r < c->world; ++r) {
    void* base = c->local_base;
    // a peer entry that carries this rank's own handle maps to the local area (single-process tests of the timeout path:
    // nobody ever writes that "peer's" records)
    if (r != rank && memcmp(handles + (size_t)r * RHSEG_XCHG_HANDLE_BYTES, c->own_handle, RHSEG_XCHG_HANDLE_BYTES) != 0) {
      cudaIpcMemHandle_t h;
      memcpy(&h, handles + (size_t)r * RHSEG_XCHG_HANDLE_BYTES, sizeof(h));
      RHSEG_CUDA(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    }
    c->peer_base[r] = base;
    c->dev.recv[r] = static_cast<uint4*>(base);
  }
  c->dev.epoch = reinterpret_cast<uint32_t*>(static_cast<unsigned char*>(c->local_base) + c->recv_bytes);
  c->dev.status = c->dev.epoch + XCHG_MAX_CTA;
  c->dev.rank = rank;
  c->dev.world = c->world;
  c->connected = true;
  return RHSEG_OK;
}

extern "C" int rhseg_xchg_all_reduce(void* ctx, const double* summary, long n_sum, const void* const* srcs,
                                     const long* counts, int n, double* out, void* stream) {
  XchgCtx* c = static_cast<XchgCtx*>(ctx);
  if (!c || !c->connected || !out || n_sum < 0 || (n_sum > 0 && !summary) || n < 0) return RHSEG_ERR_ARG;
  XchgParts p;
  p.n = 0;
  long parts_total = 0;
  if (n > 0) {
    const int rc = fill_parts(srcs, counts, n, false, p, parts_total);
    if (rc != RHSEG_OK) return rc;
  }
  const long total = n_sum + parts_total;
  if (total == 0) return RHSEG_OK;
  if (total > c->dev.cap) return RHSEG_ERR_ARG;
  // the grid is a function of `total` only (the same on every rank); all its CTAs are co-resident
  const unsigned grid = (unsigned)std::max<long>(1, std::min<long>(XCHG_MAX_CTA, (total + XCHG_THREADS * XCHG_ELEMS - 1) / (XCHG_THREADS * XCHG_ELEMS)));
  launch_pdl(xchg_all_reduce_kernel, dim3(grid), dim3(XCHG_THREADS), 0, (cudaStream_t)stream, c->dev, summary, n_sum, p, total, out);
  RHSEG_LAUNCH_CHECK();
  return RHSEG_OK;
}

extern "C" int rhseg_xchg_status(void* ctx, int* status_out) {
  XchgCtx* c = static_cast<XchgCtx*>(ctx);
  if (!c || !c->connected || !status_out) return RHSEG_ERR_ARG;
  uint32_t s = 0;
  RHSEG_CUDA(cudaMemcpy(&s, c->dev.status, sizeof(s), cudaMemcpyDeviceToHost));
  *status_out = (int)s;
  return RHSEG_OK;
}

extern "C" int rhseg_xchg_set_timeout_ms(void* ctx, long ms) {
  XchgCtx* c = static_cast<XchgCtx*>(ctx);
  if (!c || ms < 0) return RHSEG_ERR_ARG;
  c->dev.timeout_ns = (unsigned long long)ms * 1000000ull;  // read by the launches that follow (a captured graph keeps its own)
  return RHSEG_OK;
}

extern "C" int rhseg_xchg_destroy(void* ctx) {
  XchgCtx* c = static_cast<XchgCtx*>(ctx);
  if (!c) return RHSEG_ERR_ARG;
  if (c->connected)
    for (int r = 0; r < c->world; ++r)
      if (r != c->dev.rank && c->peer_base[r] && c->peer_base[r] != c->local_base) cudaIpcCloseMemHandle(c->peer_base[r]);
  if (c->local_base) cudaFree(c->local_base);
  delete c;
  return RHSEG_OK;
}

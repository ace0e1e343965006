// Per-pixel training evaluation shared by the stand-alone eval kernel (metrics.cu) and the fused
// HRNet hi-res forward (head_fwd.cu): CE/Dice statistics, bit-exact train-path prediction,
// confusion counting and consistency sums.  See rhseg_level_eval in include/rhseg_b200.h.
#pragma once
#include "common.cuh"

namespace rhseg {

// class index of one pixel following ProcessClasses (performance_metrics.py:31-47)
template <int K>
__device__ __forceinline__ int process_class(const float (&x)[K], bool child) {
  if (child) {
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < K; ++k) sum += x[k];
    float best = (sum == 0.f) ? 1.0f : 0.0f;  // prepended "nothing positive" channel
    int idx = 0;
#pragma unroll
    for (int k = 0; k < K; ++k)
      if (beats(x[k], best)) { best = x[k]; idx = k + 1; }
    return idx;
  }
  float best = x[0];
  int idx = 0;
#pragma unroll
  for (int k = 1; k < K; ++k)
    if (beats(x[k], best)) { best = x[k]; idx = k; }
  return idx;
}

// Warp-cooperative confusion counting without atomics or match: every lane owns up to SLOTS
// cells (cell = lane + 32*slot).  The warp ballots the BITS of (target class, predicted class) —
// 2*NBITS + 1 votes — and each lane ANDs the vote masks (or their complements) that spell its own
// cell, then counts the surviving lanes.  tc < 0 marks an ignored pixel.
template <int NCMAX>
struct WarpConfusion {
  static constexpr int SLOTS = (NCMAX * NCMAX + 31) / 32;
  static constexpr int NBITS = NCMAX <= 2 ? 1 : (NCMAX <= 4 ? 2 : (NCMAX <= 8 ? 3 : 4));
  int cnt[SLOTS];
  unsigned xa[SLOTS][NBITS], xb[SLOTS][NBITS];  // 0 if the cell's bit is set, ~0 otherwise (vote ^ x selects matches)
  unsigned live[SLOTS];                          // ~0 if this lane owns a real cell in the slot
  __device__ __forceinline__ void init(int nc) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      const int cell = lane + 32 * s;
      cnt[s] = 0;
      const int a = cell / nc, b = cell - a * nc;
      live[s] = cell < nc * nc ? 0xffffffffu : 0u;
#pragma unroll
      for (int i = 0; i < NBITS; ++i) {
        xa[s][i] = ((a >> i) & 1) ? 0u : 0xffffffffu;
        xb[s][i] = ((b >> i) & 1) ? 0u : 0xffffffffu;
      }
    }
  }
  __device__ __forceinline__ void add(int tc, int pc) {
    unsigned m[SLOTS];
    const unsigned valid = __ballot_sync(0xffffffffu, tc >= 0);
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) m[s] = valid & live[s];
#pragma unroll
    for (int i = 0; i < NBITS; ++i) {
      const unsigned va = __ballot_sync(0xffffffffu, (tc >> i) & 1);
      const unsigned vb = __ballot_sync(0xffffffffu, (pc >> i) & 1);
#pragma unroll
      for (int s = 0; s < SLOTS; ++s) m[s] &= (va ^ xa[s][i]) & (vb ^ xb[s][i]);
    }
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) cnt[s] += __popc(m[s]);
  }
  __device__ __forceinline__ void flush(int* hist, int nc) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int s = 0; s < SLOTS; ++s) {
      const int cell = lane + 32 * s;
      if (cell < nc * nc && cnt[s]) atomicAdd(&hist[cell], cnt[s]);
    }
  }
};

// CT: compile-time knowledge of the level kind (-1 unknown: runtime flags; 0 root level; 1 child level without
// consistency inputs; 2 child level with consistency) — the hot kernels instantiate the three kinds so that the
// per-pixel code carries no flag tests.
template <int K, int CT = -1>
struct EvalAccum {
  static constexpr int NS = RHSEG_NSTAT;
  static constexpr int NACC = K * NS + K;  // statistics + one consistency accumulator per group-start channel
  float a[K][NS], ca[K];
  WarpConfusion<K + 1> wc;
  LevelInfo li;
  int child, nc;
  bool do_cons;
  __device__ __forceinline__ bool is_child() const { return CT < 0 ? child != 0 : CT >= 1; }
  __device__ __forceinline__ bool has_cons() const { return CT < 0 ? do_cons : CT == 2; }

  // hist: shared int[(K+1)^2], zeroed here (contains a __syncthreads)
  __device__ __forceinline__ void init(int child_, bool do_cons_, const int32_t* table, int* hist, int nthreads) {
    child = child_;
    nc = child ? K + 1 : K;
    do_cons = do_cons_;
    for (int i = threadIdx.x; i < nc * nc; i += nthreads) hist[i] = 0;
    __syncthreads();
    li = load_level_info<K>(child ? table : nullptr);
    wc.init(nc);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      ca[k] = 0.f;
#pragma unroll
      for (int j = 0; j < NS; ++j) a[k][j] = 0.f;
    }
  }

  // targets of the level (+ the parent's target channel per group and the previous level's index map)
  template <int VEC>
  __device__ __forceinline__ void load_targets(const float* __restrict__ targets, long t_bs, long t_cs,
                                               const float* __restrict__ parent_targets, long pt_bs, long pt_cs,
                                               const unsigned char* __restrict__ prev_idx, int b, long N, long px,
                                               bool ok, float (&t)[K][VEC], float (&ptv)[K][VEC],
                                               unsigned char (&pidx)[VEC]) const {
#pragma unroll
    for (int v = 0; v < VEC; ++v) pidx[v] = 0;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      Vec<VEC> tv;
#pragma unroll
      for (int v = 0; v < VEC; ++v) { tv.v[v] = -1.f; ptv[k][v] = -1.f; }
      if (ok) {
        tv = ld_cached<VEC>(targets + (size_t)b * t_bs + (size_t)k * t_cs + px);
        if (has_cons() && ((li.start_mask >> k) & 1)) {
          const Vec<VEC> pv = ld_cached<VEC>(parent_targets + (size_t)b * pt_bs + (size_t)li.parent[k] * pt_cs + px);
#pragma unroll
          for (int v = 0; v < VEC; ++v) ptv[k][v] = pv.v[v];
        }
      }
#pragma unroll
      for (int v = 0; v < VEC; ++v) t[k][v] = tv.v[v];
    }
    if (ok && has_cons()) {
      if constexpr (VEC == 4) {
        const uchar4 q = *reinterpret_cast<const uchar4*>(prev_idx + (size_t)b * N + px);
        pidx[0] = q.x; pidx[1] = q.y; pidx[2] = q.z; pidx[3] = q.w;
      } else if constexpr (VEC == 2) {
        const uchar2 q = *reinterpret_cast<const uchar2*>(prev_idx + (size_t)b * N + px);
        pidx[0] = q.x; pidx[1] = q.y;
      } else {
        pidx[0] = prev_idx[(size_t)b * N + px];
      }
    }
  }

  template <int VEC>
  __device__ __forceinline__ void store_idx(unsigned char* __restrict__ idx_out, int b, long N, long px, bool ok,
                                            const unsigned char (&my_idx)[VEC]) const {
    if (ok && idx_out) {
      if constexpr (VEC == 4) {
        *reinterpret_cast<uchar4*>(idx_out + (size_t)b * N + px) = make_uchar4(my_idx[0], my_idx[1], my_idx[2], my_idx[3]);
      } else if constexpr (VEC == 2) {
        *reinterpret_cast<uchar2*>(idx_out + (size_t)b * N + px) = make_uchar2(my_idx[0], my_idx[1]);
      } else {
        idx_out[(size_t)b * N + px] = my_idx[0];
      }
    }
  }

  // one pixel (all lanes of the warp must call it: the confusion counting uses ballots); returns the
  // predicted channel index
  __device__ __forceinline__ int pixel(const float (&zz)[K], const float (&tt)[K], const float (&pt)[K], int pidx, bool ok) {
    float p[K], mx, sum;
    fast_softmax<K>(zz, p, mx, sum);
    const float lse = __logf(sum);
    const int idx = argmax_softmax_aten<K>(zz);  // train.py:219-221, bit-exact
    float pr[K], et[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      // out-of-range pixels arrive with tt = -1 (load_targets), so the mask alone gates them
      const bool m = tt[k] != -1.0f;
      const float mf = m ? 1.0f : 0.0f;
      et[k] = m ? tt[k] : 0.0f;  // eval target; also t*mask for the statistics
      const float lp = (zz[k] - mx) - lse;
      a[k][0] = fmaf(et[k], lp, a[k][0]);
      a[k][1] += mf;
      a[k][2] = fmaf(p[k], et[k], a[k][2]);
      a[k][3] = fmaf(p[k], mf, a[k][3]);
      a[k][4] += et[k];
      pr[k] = (k == idx) ? mf : 0.0f;
    }
    {
      // ProcessClasses of the masked one-hot prediction: the predicted channel if its target is not
      // ignored, otherwise "nothing positive" (class 0 on child levels, argmax of zeros = 0 else)
      float pm = 0.f;
#pragma unroll
      for (int k = 0; k < K; ++k) pm += pr[k];
      const int pc = pm != 0.f ? (is_child() ? idx + 1 : idx) : 0;
      int tc = process_class<K>(et, is_child());
      if (!ok || (is_child() && tc == 0)) tc = -1;  // out of range / torchmetrics ignore_index=0 on child levels
      wc.add(tc, pc);
    }
    if (has_cons() && ok) {
      float gs[K];
      group_sum<K>(pr, li.start_mask, gs);  // children one-hots summed per group
#pragma unroll
      for (int k = 0; k < K; ++k)
        if ((li.start_mask >> k) & 1) {
          const float parent_hot = (pidx == li.parent[k] && pt[k] != -1.0f) ? 1.0f : 0.0f;
          ca[k] += fabsf(gs[k] - parent_hot);
        }
    }
    return idx;
  }

  // block reduction + one fp64 / int64 atomic per CTA and quantity.  red: shared float[NWARP][NACC]
  template <int NWARP>
  __device__ __forceinline__ void finish(float* red, int* hist, double* __restrict__ stats, double* __restrict__ cons,
                                         unsigned long long* __restrict__ conf, const int32_t* __restrict__ table, int b) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
      for (int j = 0; j < NS; ++j) {
        const float v = warp_sum(a[k][j]);
        if (lane == 0) red[warp * NACC + k * NS + j] = v;
      }
      const float c = warp_sum(ca[k]);
      if (lane == 0) red[warp * NACC + K * NS + k] = c;
    }
    wc.flush(hist, nc);
    __syncthreads();
    if (tid < NACC) {
      double acc = 0.0;
#pragma unroll
      for (int w = 0; w < NWARP; ++w) acc += (double)red[w * NACC + tid];
      if (tid < K * NS) atomicAdd(&stats[(size_t)b * K * NS + tid], acc);
      else if (has_cons() && ((li.start_mask >> (tid - K * NS)) & 1)) atomicAdd(&cons[table[RHSEG_TBL_GROUP_OF + tid - K * NS]], acc);
    }
    for (int i = tid; i < nc * nc; i += NWARP * 32)
      if (hist[i]) atomicAdd(&conf[i], (unsigned long long)hist[i]);
  }
};

}  // namespace rhseg

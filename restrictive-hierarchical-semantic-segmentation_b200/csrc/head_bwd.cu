// Head backward: activation backward (composition / restrictive softmax / sigmoid), adjoint
// of the align_corners bilinear upsample, 1x1-conv backward (dfeats + fp64 weight-gradient
// sums) and the tiny parameter-gradient kernel (head conv, FiLM linear, uniform gradient to
// the previous level's pooled probabilities).  The reference obtains all of this from
// autograd over Models/models.py:58-77 and :263-306 / :757-802; closed forms in DESIGN.md.
#include <algorithm>
#include "common.cuh"

namespace rhseg {

// ------------------------------------------------------------------------------------
// activation backward, output resolution, elementwise
// ------------------------------------------------------------------------------------
template <int K, int VEC, int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS)
act_bwd_kernel(const float* __restrict__ logits, const float* __restrict__ prev_probs,
               const int32_t* __restrict__ table, const float* __restrict__ dz_in,
               const double* __restrict__ g_uniform, float inv_npix, const float* __restrict__ dp_pix,
               uint32_t pix_mask, int K_prev, long N, float* __restrict__ dz_out, float* __restrict__ dp_prev) {
  const int b = blockIdx.y;
  const long px = ((long)blockIdx.x * THREADS + threadIdx.x) * VEC;
  if (px >= N) return;
  const LevelInfo li = load_level_info<K>(MODE == RHSEG_ACT_GROUPED ? table : nullptr);
  const size_t base = (size_t)b * K * N + px;

  float z[K][VEC], dP[K][VEC], dz[K][VEC];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const Vec<VEC> t = ld_stream<VEC>(logits + base + (size_t)k * N);
    const float gu = g_uniform ? (float)(g_uniform[b * K + k]) * inv_npix : 0.f;
    Vec<VEC> d, e;
#pragma unroll
    for (int v = 0; v < VEC; ++v) { d.v[v] = 0.f; e.v[v] = 0.f; }
    if (dz_in) d = ld_stream<VEC>(dz_in + base + (size_t)k * N);
    if (dp_pix && ((pix_mask >> k) & 1u)) e = ld_stream<VEC>(dp_pix + base + (size_t)k * N);
#pragma unroll
    for (int v = 0; v < VEC; ++v) { z[k][v] = t.v[v]; dz[k][v] = d.v[v]; dP[k][v] = gu + e.v[v]; }
  }

  if constexpr (MODE == RHSEG_ACT_SIGMOID) {
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const float p = sigmoidf_ref(z[k][v]);
        dz[k][v] = fmaf(dP[k][v], p * (1.0f - p), dz[k][v]);
      }
  } else if constexpr (MODE == RHSEG_ACT_GROUPED) {
    float pp[K][VEC];
    const float* pb = prev_probs + (size_t)b * K_prev * N + px;
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if ((li.start_mask >> k) & 1) {
        const Vec<VEC> t = ld_stream<VEC>(pb + (size_t)li.parent[k] * N);
#pragma unroll
        for (int v = 0; v < VEC; ++v) pp[k][v] = t.v[v];
      } else {
#pragma unroll
        for (int v = 0; v < VEC; ++v) pp[k][v] = pp[k > 0 ? k - 1 : 0][v];
      }
    }
    float dpar[K][VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      float zz[K], q[K], dq_q[K], inner[K], dpq[K], dparent[K];
#pragma unroll
      for (int k = 0; k < K; ++k) zz[k] = z[k][v];
      grouped_softmax<K>(zz, li.start_mask, q);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        dpq[k] = dP[k][v] * q[k];          // P_c = P_p * Q_c  ->  dL/dP_p += dP_c * Q_c
        dq_q[k] = dpq[k] * pp[k][v];       // dQ_c * Q_c with dQ_c = dP_c * P_p
      }
      group_sum<K>(dq_q, li.start_mask, inner);
      group_sum<K>(dpq, li.start_mask, dparent);
#pragma unroll
      for (int k = 0; k < K; ++k) {
        // softmax backward inside the group; the log(P_p+eps) gate contributes exactly 0
        dz[k][v] += dq_q[k] - q[k] * inner[k];
        dpar[k][v] = dparent[k];
      }
    }
    if (dp_prev) {
      float* dpb = dp_prev + (size_t)b * K_prev * N + px;
#pragma unroll
      for (int k = 0; k < K; ++k)
        if ((li.start_mask >> k) & 1) {
          float* dst = dpb + (size_t)li.parent[k] * N;
          Vec<VEC> cur = *reinterpret_cast<const Vec<VEC>*>(dst);
#pragma unroll
          for (int v = 0; v < VEC; ++v) cur.v[v] += dpar[k][v];
          *reinterpret_cast<Vec<VEC>*>(dst) = cur;
        }
    }
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    Vec<VEC> o;
#pragma unroll
    for (int v = 0; v < VEC; ++v) o.v[v] = dz[k][v];
    *reinterpret_cast<Vec<VEC>*>(dz_out + base + (size_t)k * N) = o;
  }
}

// ------------------------------------------------------------------------------------
// adjoint of the bilinear upsample (gather form, deterministic)
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float lerp_weight(int dst, float scale, int in_size, int want) {
  const float src = scale * (float)dst;
  const int i0 = (int)src;
  const int i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  const float l1 = src - (float)i0, l0 = 1.0f - l1;
  return (i0 == want ? l0 : 0.f) + (i1 == want ? l1 : 0.f);
}
__device__ __forceinline__ void support(int i, float scale, int out_size, int& lo, int& hi) {
  if (scale <= 0.f) { lo = 0; hi = out_size - 1; return; }
  lo = max(0, (int)floorf((float)(i - 1) / scale) - 1);
  hi = min(out_size - 1, (int)ceilf((float)(i + 1) / scale) + 1);
}

// One thread per low-res element.  The (few) non-zero column weights are hoisted into
// registers; the row loop then is pure FMA over L1/L2-resident hi-res gradients.
constexpr int ADJ_MAXS = 12;  // covers upsampling factors up to ~5x; larger factors take the generic loop

__global__ void __launch_bounds__(128)
upsample_adjoint_kernel(const float* __restrict__ dz_hi, int Hf, int Wf, int H, int W, float sy, float sx,
                        long total, float* __restrict__ dz_lo) {
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int j = (int)(idx % Wf);
  const int i = (int)((idx / Wf) % Hf);
  const long bk = idx / ((long)Wf * Hf);
  int y0, y1, x0, x1;
  support(i, sy, H, y0, y1);
  support(j, sx, W, x0, x1);
  const float* src = dz_hi + (size_t)bk * H * W;
  float acc = 0.f;
  if (x1 - x0 + 1 <= ADJ_MAXS) {
    float wx[ADJ_MAXS];
#pragma unroll
    for (int t = 0; t < ADJ_MAXS; ++t) wx[t] = (x0 + t <= x1) ? lerp_weight(x0 + t, sx, Wf, j) : 0.f;
    for (int y = y0; y <= y1; ++y) {
      const float wy = lerp_weight(y, sy, Hf, i);
      if (wy == 0.f) continue;
      const float* rowp = src + (size_t)y * W + x0;
      float row = 0.f;
#pragma unroll
      for (int t = 0; t < ADJ_MAXS; ++t)
        if (wx[t] != 0.f) row = fmaf(wx[t], __ldg(rowp + t), row);
      acc = fmaf(wy, row, acc);
    }
  } else {
    for (int y = y0; y <= y1; ++y) {
      const float wy = lerp_weight(y, sy, Hf, i);
      if (wy == 0.f) continue;
      float row = 0.f;
      for (int x = x0; x <= x1; ++x) {
        const float wx = lerp_weight(x, sx, Wf, j);
        if (wx != 0.f) row = fmaf(wx, __ldg(src + (size_t)y * W + x), row);
      }
      acc = fmaf(wy, row, acc);
    }
  }
  dz_lo[idx] = acc;
}

// ------------------------------------------------------------------------------------
// 1x1 conv backward at feature resolution, persistent CTAs.
// Work space = (sample, channel slice, pixel-vector), flattened in that order and split EVENLY
// over one resident wave of CTAs.  Each thread keeps dz for its J*VEC pixels in registers and
// walks its segment's channel slice:
//   dfeats[c] = sum_k w[k][c] dz[k]                         (pure write stream)
//   S[k][c]  += sum_pixels dz[k] * feats[c]                 (read stream + reduction)
// The per-channel K partial sums are reduced across the warp with the transposed
// recursive-halving shuffle pattern (KP-1 + 5-log2(KP) shuffles), accumulated per warp in
// shared memory over the CTA's tiles, and leave the CTA as one fp64 atomic per (k, c) per
// (sample, slice) segment.
// ------------------------------------------------------------------------------------
template <int K, int VEC, int J, int THREADS, int UNROLL>
__global__ void __launch_bounds__(THREADS)
conv_bwd_kernel(const float* __restrict__ feats, const float* __restrict__ dz, const float* __restrict__ eff_w,
                int C, int c_per_slice, int slices, int N, long total_work, float* __restrict__ dfeats,
                double* __restrict__ S, double* __restrict__ s) {
  constexpr int KP = pad_k(K);
  constexpr int P = J * VEC;
  constexpr int NWARP = THREADS / 32;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* w_t = smem;                               // [c_per_slice][KP]
  float* red = smem + (size_t)c_per_slice * KP;    // [NWARP][c_per_slice][KP]
  float* my_red = red + (size_t)warp * c_per_slice * KP;
  const long ups = N / VEC;
  const long w_begin = total_work * blockIdx.x / gridDim.x;
  const long w_end = total_work * (blockIdx.x + 1) / gridDim.x;
  const bool writer = transposed_writer<KP>(lane);
  const int my_k = transposed_index<KP>(lane);

  long cur_seg = -1;
  int b = 0, c_begin = 0, c_cnt = 0;
  float s_acc = 0.f;  // writer lanes: running sum_n dz[my_k] for the current segment (slice 0 only)

  auto flush = [&]() {  // uniform: called by the whole CTA
    __syncthreads();
    for (int i = tid; i < c_cnt * K; i += THREADS) {
      const int k = i / c_cnt, c = i - k * c_cnt;
      double acc = 0.0;
#pragma unroll
      for (int w = 0; w < NWARP; ++w) acc += (double)red[((size_t)w * c_per_slice + c) * KP + k];
      atomicAdd(&S[((size_t)b * K + k) * C + c_begin + c], acc);
    }
    if (c_begin == 0 && writer && my_k < K) atomicAdd(&s[b * K + my_k], (double)s_acc);
    __syncthreads();
  };

  long w0 = w_begin;
  while (w0 < w_end) {
    const long seg = w0 / ups;  // = b * slices + slice
    const long seg_end = min(w_end, (seg + 1) * ups);
    const long tile_end = min(seg_end, w0 + (long)THREADS * J);
    if (seg != cur_seg) {
      if (cur_seg >= 0) flush();
      b = (int)(seg / slices);
      c_begin = (int)(seg % slices) * c_per_slice;
      c_cnt = min(c_per_slice, C - c_begin);
      for (int i = tid; i < K * c_cnt; i += THREADS) {
        const int k = i / c_cnt, c = i - k * c_cnt;
        w_t[c * KP + k] = eff_w[((size_t)b * K + k) * C + c_begin + c];
      }
      if constexpr (KP > K)
        for (int i = tid; i < (KP - K) * c_cnt; i += THREADS) w_t[(i % c_cnt) * KP + K + i / c_cnt] = 0.f;
      for (int i = tid; i < NWARP * c_per_slice * KP; i += THREADS) red[i] = 0.f;
      s_acc = 0.f;
      cur_seg = seg;
      __syncthreads();
    }

    long px[J];
    bool ok[J];
    float g[K][P];
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const long u = w0 + (long)j * THREADS + tid;
      ok[j] = u < tile_end;
      px[j] = (u - seg * ups) * VEC;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        Vec<VEC> t;
        if (ok[j]) t = ld_cached<VEC>(dz + ((size_t)b * K + k) * N + px[j]);
        else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) t.v[v] = 0.f;
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) g[k][j * VEC + v] = t.v[v];
      }
    }

    const float* fb = feats + ((size_t)b * C + c_begin) * N;
    float* dfb = dfeats ? dfeats + ((size_t)b * C + c_begin) * N : nullptr;

    auto one_channel = [&](int cl, const Vec<VEC> (&f)[J]) {
      float w[KP];
#pragma unroll
      for (int k = 0; k < KP; ++k) w[k] = w_t[cl * KP + k];
      if (dfb) {
#pragma unroll
        for (int j = 0; j < J; ++j) {
          if (!ok[j]) continue;
          Vec<VEC> o;
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < K; ++k) a = fmaf(w[k], g[k][j * VEC + v], a);
            o.v[v] = a;
          }
          st_stream<VEC>(dfb + (size_t)cl * N + px[j], o);
        }
      }
      float part[KP];
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        part[k] = 0.f;
        if (k < K) {
#pragma unroll
          for (int j = 0; j < J; ++j)
#pragma unroll
            for (int v = 0; v < VEC; ++v) part[k] = fmaf(g[k][j * VEC + v], f[j].v[v], part[k]);
        }
      }
      warp_reduce_transposed<KP>(part, lane);
      if (writer) my_red[cl * KP + my_k] += part[0];
    };

    int cl = 0;
    for (; cl + UNROLL <= c_cnt; cl += UNROLL) {
      Vec<VEC> f[UNROLL][J];
#pragma unroll
      for (int u = 0; u < UNROLL; ++u)
#pragma unroll
        for (int j = 0; j < J; ++j) {
          if (ok[j]) f[u][j] = ld_stream<VEC>(fb + (size_t)(cl + u) * N + px[j]);
          else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) f[u][j].v[v] = 0.f;
          }
        }
#pragma unroll
      for (int u = 0; u < UNROLL; ++u) one_channel(cl + u, f[u]);
    }
    for (; cl < c_cnt; ++cl) {
      Vec<VEC> f[J];
#pragma unroll
      for (int j = 0; j < J; ++j) {
        if (ok[j]) f[j] = ld_stream<VEC>(fb + (size_t)cl * N + px[j]);
        else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) f[j].v[v] = 0.f;
        }
      }
      one_channel(cl, f);
    }

    if (c_begin == 0) {  // s[b][k] = sum_n dz: only the first channel slice contributes
      float sred[KP];
#pragma unroll
      for (int k = 0; k < KP; ++k) {
        sred[k] = 0.f;
        if (k < K) {
#pragma unroll
          for (int p = 0; p < P; ++p) sred[k] += g[k][p];
        }
      }
      warp_reduce_transposed<KP>(sred, lane);
      s_acc += sred[0];
    }
    w0 = tile_end;
  }
  if (cur_seg >= 0) flush();
}

template <int K, int VEC, int J, int THREADS, int UNROLL>
static int launch_conv_bwd(const float* feats, const float* dz, const float* eff_w, int B, int C, int N,
                           float* dfeats, double* S, double* s, int sm_count, cudaStream_t st) {
  constexpr int KP = pad_k(K);
  auto kern = conv_bwd_kernel<K, VEC, J, THREADS, UNROLL>;
  const long ups = N / VEC;
  const long tile = (long)THREADS * J;
  // channel slices: enough work for >= ~4 tiles per resident CTA, slices no shorter than 2*UNROLL channels
  int slices = 1;
  auto smem_for = [&](int sl) { return (size_t)((C + sl - 1) / sl) * KP * (1 + THREADS / 32) * sizeof(float); };
  auto grid_for = [&](int sl, int* per_sm_out) {
    int per_sm = 0;
    const size_t smem = smem_for(sl);
    if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem);
    if (per_sm < 1) per_sm = 1;
    if (per_sm_out) *per_sm_out = per_sm;
    return (long)sm_count * per_sm;
  };
  while ((long)B * ups * slices < grid_for(slices, nullptr) * tile * 4 && (C + slices * 2 - 1) / (slices * 2) >= 2 * UNROLL)
    slices *= 2;
  const int c_per_slice = (C + slices - 1) / slices;
  slices = (C + c_per_slice - 1) / c_per_slice;
  const size_t smem = (size_t)c_per_slice * KP * (1 + THREADS / 32) * sizeof(float);
  if (smem > 48 * 1024) RHSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  RHSEG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
  if (per_sm < 1) per_sm = 1;
  const long total_work = (long)B * slices * ups;
  const long tiles = (total_work + tile - 1) / tile;
  const long grid = std::min<long>((long)sm_count * per_sm, tiles);
  kern<<<(unsigned)grid, THREADS, smem, st>>>(feats, dz, eff_w, C, c_per_slice, slices, N, total_work, dfeats, S, s);
  RHSEG_LAUNCH_CHECK();
  return RHSEG_OK;
}

// ------------------------------------------------------------------------------------
// parameter gradients (tiny): one thread per feature channel
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
param_grads_kernel(const double* __restrict__ S, const double* __restrict__ s, const float* __restrict__ head_w,
                   const float* __restrict__ film_w, const float* __restrict__ gamma_beta,
                   const double* __restrict__ prev_psum, double n_pix, int B, int C, int K, int K_prev,
                   float* __restrict__ d_head_w, float* __restrict__ d_head_b, float* __restrict__ d_film_w,
                   float* __restrict__ d_film_b, double* __restrict__ g_prev) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  const bool ok = c < C;
  const int lane = threadIdx.x & 31;
  if (blockIdx.x == 0 && threadIdx.x < K) {
    double acc = 0.0;
    for (int b = 0; b < B; ++b) acc += s[b * K + threadIdx.x];
    d_head_b[threadIdx.x] = (float)acc;
  }
  double dw[RHSEG_KERNEL_MAX_K];
  for (int k = 0; k < RHSEG_KERNEL_MAX_K; ++k) dw[k] = 0.0;
  double dfw_g[RHSEG_MAX_K], dfw_b[RHSEG_MAX_K], dfb_g = 0.0, dfb_b = 0.0;
  for (int j = 0; j < RHSEG_MAX_K; ++j) { dfw_g[j] = 0.0; dfw_b[j] = 0.0; }
  for (int b = 0; b < B; ++b) {
    double dgam = 0.0, dbet = 0.0;
    if (ok) {
      const double gam = film_w ? (double)gamma_beta[(size_t)b * 2 * C + c] : 1.0;
      const double bet = film_w ? (double)gamma_beta[(size_t)b * 2 * C + C + c] : 0.0;
      for (int k = 0; k < K; ++k) {
        const double Sv = S[((size_t)b * K + k) * C + c], sv = s[b * K + k];
        const double w = (double)head_w[(size_t)k * C + c];
        dw[k] += Sv * gam + sv * bet;
        dgam += Sv * w;
        dbet += w * sv;
      }
    }
    if (film_w) {
      dfb_g += dgam;
      dfb_b += dbet;
      for (int j = 0; j < K_prev; ++j) {
        const double cond = (double)(float)(prev_psum[b * K_prev + j] / n_pix);
        dfw_g[j] += dgam * cond;
        dfw_b[j] += dbet * cond;
        double contrib = 0.0;
        if (ok) contrib = (double)film_w[(size_t)c * K_prev + j] * dgam + (double)film_w[(size_t)(C + c) * K_prev + j] * dbet;
        contrib = warp_sum(contrib);
        if (lane == 0) atomicAdd(&g_prev[b * K_prev + j], contrib);
      }
    }
  }
  if (!ok) return;
  for (int k = 0; k < K; ++k) d_head_w[(size_t)k * C + c] = (float)dw[k];
  if (film_w) {
    d_film_b[c] = (float)dfb_g;
    d_film_b[C + c] = (float)dfb_b;
    for (int j = 0; j < K_prev; ++j) {
      d_film_w[(size_t)c * K_prev + j] = (float)dfw_g[j];
      d_film_w[(size_t)(C + c) * K_prev + j] = (float)dfw_b[j];
    }
  }
}

}  // namespace rhseg

using namespace rhseg;

extern "C" int rhseg_head_act_bwd(const float* logits, const float* prev_probs, const int32_t* table,
                                  const float* dz_in, const double* g_uniform, double inv_npix,
                                  const float* dp_pix, uint32_t pix_mask, int B, int K, int K_prev, int H, int W,
                                  int act_mode, float* dz_out, float* dp_prev, void* stream) {
  if (!logits || !dz_out || B <= 0 || H <= 0 || W <= 0) return RHSEG_ERR_ARG;
  if (K < 1 || K > RHSEG_KERNEL_MAX_K) return RHSEG_ERR_UNSUPPORTED;
  if (act_mode == RHSEG_ACT_GROUPED && (!prev_probs || !table)) return RHSEG_ERR_ARG;
  if (dz_out == dz_in) return RHSEG_ERR_ARG;
  const long N = (long)H * W;
  cudaStream_t st = (cudaStream_t)stream;
  constexpr int THREADS = 256;
  const float inv = (float)inv_npix;
  RHSEG_DISPATCH_K(K, {
    if (N % 4 == 0) {
      dim3 grid((unsigned)((N / 4 + THREADS - 1) / THREADS), B);
      if (act_mode == RHSEG_ACT_SIGMOID)
        act_bwd_kernel<KK, 4, RHSEG_ACT_SIGMOID, THREADS><<<grid, THREADS, 0, st>>>(logits, prev_probs, table, dz_in, g_uniform, inv, dp_pix, pix_mask, K_prev, N, dz_out, dp_prev);
      else if (act_mode == RHSEG_ACT_GROUPED)
        act_bwd_kernel<KK, 4, RHSEG_ACT_GROUPED, THREADS><<<grid, THREADS, 0, st>>>(logits, prev_probs, table, dz_in, g_uniform, inv, dp_pix, pix_mask, K_prev, N, dz_out, dp_prev);
      else
        act_bwd_kernel<KK, 4, RHSEG_ACT_ZEROS, THREADS><<<grid, THREADS, 0, st>>>(logits, prev_probs, table, dz_in, g_uniform, inv, dp_pix, pix_mask, K_prev, N, dz_out, dp_prev);
    } else {
      dim3 grid((unsigned)((N + THREADS - 1) / THREADS), B);
      if (act_mode == RHSEG_ACT_SIGMOID)
        act_bwd_kernel<KK, 1, RHSEG_ACT_SIGMOID, THREADS><<<grid, THREADS, 0, st>>>(logits, prev_probs, table, dz_in, g_uniform, inv, dp_pix, pix_mask, K_prev, N, dz_out, dp_prev);
      else if (act_mode == RHSEG_ACT_GROUPED)
        act_bwd_kernel<KK, 1, RHSEG_ACT_GROUPED, THREADS><<<grid, THREADS, 0, st>>>(logits, prev_probs, table, dz_in, g_uniform, inv, dp_pix, pix_mask, K_prev, N, dz_out, dp_prev);
      else
        act_bwd_kernel<KK, 1, RHSEG_ACT_ZEROS, THREADS><<<grid, THREADS, 0, st>>>(logits, prev_probs, table, dz_in, g_uniform, inv, dp_pix, pix_mask, K_prev, N, dz_out, dp_prev);
    }
  });
  RHSEG_LAUNCH_CHECK();
  return RHSEG_OK;
}

extern "C" int rhseg_upsample_adjoint(const float* dz_hi, int BK, int Hf, int Wf, int H, int W, float* dz_lo,
                                      void* stream) {
  if (!dz_hi || !dz_lo || BK <= 0 || Hf <= 0 || Wf <= 0 || H <= 0 || W <= 0) return RHSEG_ERR_ARG;
  const float sy = H > 1 ? (float)(Hf - 1) / (float)(H - 1) : 0.f;
  const float sx = W > 1 ? (float)(Wf - 1) / (float)(W - 1) : 0.f;
  const long total = (long)BK * Hf * Wf;
  upsample_adjoint_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(dz_hi, Hf, Wf, H, W, sy, sx,
                                                                                           total, dz_lo);
  RHSEG_LAUNCH_CHECK();
  return RHSEG_OK;
}

extern "C" int rhseg_head_conv_bwd(const float* feats, const float* dz, const float* eff_w, int B, int C, int K,
                                   int n_pix, float* dfeats, double* S, double* s, int zero_sums, void* stream) {
  if (!feats || !dz || !eff_w || !S || !s || B <= 0 || C <= 0 || n_pix <= 0) return RHSEG_ERR_ARG;
  if (K < 1 || K > RHSEG_KERNEL_MAX_K) return RHSEG_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  if (zero_sums) {
    RHSEG_CUDA(cudaMemsetAsync(S, 0, sizeof(double) * (size_t)B * K * C, st));
    RHSEG_CUDA(cudaMemsetAsync(s, 0, sizeof(double) * (size_t)B * K, st));
  }
  const int sms = device_sm_count();
  RHSEG_DISPATCH_K(K, {
    if (n_pix % 4 == 0) return launch_conv_bwd<KK, 4, (KK <= 4 ? 2 : 1), 256, 4>(feats, dz, eff_w, B, C, n_pix, dfeats, S, s, sms, st);
    return launch_conv_bwd<KK, 1, (KK <= 4 ? 4 : 2), 128, 4>(feats, dz, eff_w, B, C, n_pix, dfeats, S, s, sms, st);
  });
  return RHSEG_OK;
}

extern "C" int rhseg_head_param_grads(const double* S, const double* s, const float* head_w, const float* film_w,
                                      const float* gamma_beta, const double* prev_psum, double n_pix, int B, int C,
                                      int K, int K_prev, float* d_head_w, float* d_head_b, float* d_film_w,
                                      float* d_film_b, double* g_prev, void* stream) {
  if (!S || !s || !head_w || !d_head_w || !d_head_b || B <= 0 || C <= 0) return RHSEG_ERR_ARG;
  if (K < 1 || K > RHSEG_KERNEL_MAX_K) return RHSEG_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  if (film_w) {
    if (!gamma_beta || !prev_psum || !d_film_w || !d_film_b || !g_prev || n_pix <= 0) return RHSEG_ERR_ARG;
    if (K_prev < 1 || K_prev > RHSEG_MAX_K) return RHSEG_ERR_UNSUPPORTED;
    RHSEG_CUDA(cudaMemsetAsync(g_prev, 0, sizeof(double) * (size_t)B * K_prev, st));
  }
  param_grads_kernel<<<(C + 127) / 128, 128, 0, st>>>(S, s, head_w, film_w, gamma_beta, prev_psum, n_pix, B, C, K, K_prev,
                                                      d_head_w, d_head_b, d_film_w, d_film_b, g_prev);
  RHSEG_LAUNCH_CHECK();
  return RHSEG_OK;
}

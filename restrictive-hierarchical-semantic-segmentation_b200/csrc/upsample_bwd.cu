// Adjoint of the bilinear (align_corners=True) upsample, shared-memory tiled and separable,
// optionally fused with the hi-res gradient producers (HRNet path).
//
// The reference reaches this through autograd over F.interpolate (Models/models.py:766, :776):
//   dz_lo[i][j] = sum_{y,x} wy(y,i) wx(x,j) dz_hi[y][x]
// One CTA owns a tile of TH x TW low-res pixels of one sample (all K channels).  It stages the
// hi-res gradient of the tile's support region in shared memory, reduces along x, then along y.
// SRC 0: dz_hi is read from memory (what autograd hands us).
// SRC 1: dz_hi is formed on the fly per hi-res pixel as  loss gradient (closed form from logits,
//        targets, coefficients) + activation backward (sigmoid / restrictive softmax with the
//        gradient arriving at the probabilities), so the hi-res gradient tensor never exists.
#include <algorithm>
#include "common.cuh"

namespace rhseg {

constexpr int ADJ_TH = 8, ADJ_TW = 16, ADJ_THREADS = 256;
constexpr int ADJ_MAXW = 24;  // max hi-res taps per low-res index held in the weight tables (upsampling factor <= ~10)

struct FusedDzArgs {
  const float* logits;      // [B,K,H,W]
  const float* targets;     // strided
  long t_bstride, t_cstride;
  const float* coef;        // [B,K,3]
  const float* g_ce;        // device scalars (may be null)
  const float* g_dice;
  const float* prev_probs;  // [B,K_prev,H,W] (grouped)
  const int32_t* table;
  const double* g_uniform;  // [B,K] or null
  float inv_npix;
  const float* dp_pix;      // [B,K,H,W] or null
  uint32_t pix_mask;
  float* dp_prev;           // [B,K_prev,H,W] (+=) or null
  int K_prev;
};

template <int K, int SRC, int MODE>
__global__ void __launch_bounds__(ADJ_THREADS)
upsample_adjoint_tiled_kernel(const float* __restrict__ dz_hi, FusedDzArgs fa, int Hf, int Wf, int H, int W, float sy,
                              float sx, int ry_max, int rx_max, float* __restrict__ dz_lo) {
  pdl_wait();
  extern __shared__ __align__(16) float sm[];
  float* reg = sm;                                  // [K][ry_max][rx_max]   hi-res gradient of the support region
  float* tmp = sm + (size_t)K * ry_max * rx_max;    // [K][ry_max][ADJ_TW]   after the x reduction
  const int b = blockIdx.z, tid = threadIdx.x;
  const int i0 = blockIdx.y * ADJ_TH, j0 = blockIdx.x * ADJ_TW;
  const int th = min(ADJ_TH, Hf - i0), tw = min(ADJ_TW, Wf - j0);
  int y0, y1, x0, x1, dummy;
  lerp_support(i0, sy, H, y0, dummy);
  lerp_support(i0 + th - 1, sy, H, dummy, y1);
  lerp_support(j0, sx, W, x0, dummy);
  lerp_support(j0 + tw - 1, sx, W, dummy, x1);
  const int ry = y1 - y0 + 1, rx = x1 - x0 + 1;  // <= ry_max, rx_max by construction of the launcher
  const long N = (long)H * W;

  // interpolation weight tables of this tile: for output column j0+tj the contiguous run of hi-res
  // columns [wx_start, wx_start+wx_cnt) that read it, with their weights (same for rows)
  __shared__ float wx_tab[ADJ_TW * ADJ_MAXW], wy_tab[ADJ_TH * ADJ_MAXW];
  __shared__ int wx_start[ADJ_TW], wx_cnt[ADJ_TW], wy_start[ADJ_TH], wy_cnt[ADJ_TH];
  if (tid < ADJ_TW + ADJ_TH) {
    const bool is_x = tid < ADJ_TW;
    const int t = is_x ? tid : tid - ADJ_TW;
    const int lim = is_x ? tw : th;
    if (t < lim) {
      const int want = (is_x ? j0 : i0) + t;
      const float sc = is_x ? sx : sy;
      const int in_size = is_x ? Wf : Hf, out_size = is_x ? W : H;
      int lo, hi;
      lerp_support(want, sc, out_size, lo, hi);
      int first = -1, cnt = 0;
      float* tab = (is_x ? wx_tab : wy_tab) + t * ADJ_MAXW;
      for (int o = lo; o <= hi; ++o) {
        const float w = lerp_weight(o, sc, in_size, want);
        if (w != 0.f || first >= 0) {
          if (first < 0) first = o;
          if (cnt < ADJ_MAXW) tab[cnt] = w;
          ++cnt;
        }
      }
      // trailing zeros are harmless; the run is contiguous because the weights form a hat function
      (is_x ? wx_start : wy_start)[t] = first < 0 ? lo : first;
      (is_x ? wx_cnt : wy_cnt)[t] = cnt < ADJ_MAXW ? cnt : ADJ_MAXW;
    }
  }

  // ---- phase 1: hi-res gradient of the region -> shared memory ----
  if constexpr (SRC == 0) {
    const int lane = tid & 31, warp = tid >> 5;
    for (int yy = warp; yy < ry; yy += ADJ_THREADS / 32)
      for (int xx = lane; xx < rx; xx += 32) {
        const float* src = dz_hi + (size_t)b * K * N + (size_t)(y0 + yy) * W + x0 + xx;
#pragma unroll
        for (int k = 0; k < K; ++k) reg[((size_t)k * ry_max + yy) * rx_max + xx] = __ldg(src + (size_t)k * N);
      }
  } else {
    const LevelInfo li = load_level_info<K>(MODE == RHSEG_ACT_GROUPED ? fa.table : nullptr);
    const float gce = fa.g_ce ? __ldg(fa.g_ce) : 0.f, gdi = fa.g_dice ? __ldg(fa.g_dice) : 0.f;
    float A[K], Bc[K], Cc[K], gu[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float* cf = fa.coef + ((size_t)b * K + k) * 3;
      A[k] = gce * __ldg(cf);
      Bc[k] = gdi * __ldg(cf + 1);
      Cc[k] = gdi * __ldg(cf + 2);
      gu[k] = fa.g_uniform ? (float)fa.g_uniform[b * K + k] * fa.inv_npix : 0.f;
    }
    const bool has_act = MODE != RHSEG_ACT_ZEROS && (fa.g_uniform != nullptr || (fa.dp_pix != nullptr && fa.pix_mask != 0));
    const int lane = tid & 31, warp = tid >> 5;
    for (int yy = warp; yy < ry; yy += ADJ_THREADS / 32)
     for (int xx = lane; xx < rx; xx += 32) {
      const int y = y0 + yy, x = x0 + xx;
      const size_t px = (size_t)y * W + x;
      float z[K], t[K], dz[K];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        z[k] = __ldg(fa.logits + ((size_t)b * K + k) * N + px);
        t[k] = __ldg(fa.targets + (size_t)b * fa.t_bstride + (size_t)k * fa.t_cstride + px);
      }
      loss_dz_pixel<K, true>(z, t, A, Bc, Cc, dz);
      if (has_act) {
        float dP[K], pp[K], dpar[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
          dP[k] = gu[k];
          if (fa.dp_pix && ((fa.pix_mask >> k) & 1u)) dP[k] += __ldg(fa.dp_pix + ((size_t)b * K + k) * N + px);
          pp[k] = 0.f;
        }
        if constexpr (MODE == RHSEG_ACT_GROUPED) {
#pragma unroll
          for (int k = 0; k < K; ++k)
            pp[k] = ((li.start_mask >> k) & 1) ? __ldg(fa.prev_probs + ((size_t)b * fa.K_prev + li.parent[k]) * N + px)
                                               : pp[k > 0 ? k - 1 : 0];
        }
        act_dz_pixel<K, MODE>(z, dP, pp, li.start_mask, dz, dpar);
        if constexpr (MODE == RHSEG_ACT_GROUPED) {
          // every hi-res pixel is owned by exactly one tile: the one holding floor(src index)
          if (fa.dp_prev) {
            const int oi = (int)(sy * (float)y), oj = (int)(sx * (float)x);
            if (oi >= i0 && oi < i0 + th && oj >= j0 && oj < j0 + tw) {
#pragma unroll
              for (int k = 0; k < K; ++k)
                if ((li.start_mask >> k) & 1) fa.dp_prev[((size_t)b * fa.K_prev + li.parent[k]) * N + px] += dpar[k];
            }
          }
        }
      }
#pragma unroll
      for (int k = 0; k < K; ++k) reg[((size_t)k * ry_max + yy) * rx_max + xx] = dz[k];
    }
  }
  __syncthreads();

  // ---- phase 2: reduce along x: tmp[k][yy][tj] = sum_x wx(x, j0+tj) reg[k][yy][x] ----
  // (the few non-zero interpolation weights per output column / row were tabulated once per CTA)
  {
    const int tj = tid & (ADJ_TW - 1), rslot = tid / ADJ_TW;
    if (tj < tw) {
      const int xs = wx_start[tj] - x0, xn = wx_cnt[tj];
      const float* wt = wx_tab + tj * ADJ_MAXW;
#pragma unroll
      for (int k = 0; k < K; ++k)
        for (int yy = rslot; yy < ry; yy += ADJ_THREADS / ADJ_TW) {
          const float* row = reg + ((size_t)k * ry_max + yy) * rx_max + xs;
          float acc = 0.f;
          for (int q = 0; q < xn; ++q) acc = fmaf(wt[q], row[q], acc);
          tmp[((size_t)k * ry_max + yy) * ADJ_TW + tj] = acc;
        }
    }
  }
  __syncthreads();

  // ---- phase 3: reduce along y and store ----
  {
    const int tj = tid & (ADJ_TW - 1), ti = (tid / ADJ_TW) & (ADJ_TH - 1), kslot = tid / (ADJ_TW * ADJ_TH);
    if (tj < tw && ti < th) {
      const int ys = wy_start[ti] - y0, yn = wy_cnt[ti];
      const float* wt = wy_tab + ti * ADJ_MAXW;
      for (int k = kslot; k < K; k += ADJ_THREADS / (ADJ_TW * ADJ_TH)) {
        const float* col = tmp + ((size_t)k * ry_max + ys) * ADJ_TW + tj;
        float acc = 0.f;
        for (int q = 0; q < yn; ++q) acc = fmaf(wt[q], col[q * ADJ_TW], acc);
        dz_lo[((size_t)b * K + k) * Hf * Wf + (size_t)(i0 + ti) * Wf + j0 + tj] = acc;
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// Halo-free separable adjoint for row-aligned inputs (the HRNet case): pass 1 handles ROWS full
// hi-res rows per CTA (every hi-res pixel's gradient is formed exactly once), reduces them along x
// into tmpx [B,K,H,Wf]; pass 2 reduces tmpx along y.  Used when W % 4 == 0 and the taps per low-res
// index fit XR_MAXW; the tiled kernel above is the generic fallback.
// ------------------------------------------------------------------------------------
constexpr int XR_MAXROWS = 8, XR_THREADS = 256, XR_MAXW = 12;

// BAND: the CTA owns `band` consecutive hi-res rows; each x-reduced row is scattered straight into a
// shared-memory accumulator of the few low-res rows the band touches (adjoint of the y interpolation: row y
// feeds rows i0(y), i1(y) with the forward's own lerp weights) and the accumulator is added into a
// PRE-ZEROED dz_lo at the end.  band >= taps per low-res row, so a low-res row receives from at most two
// CTAs and the two-term fp32 sum is order independent: deterministic, no tmpx tensor, no second kernel.
// ACTK: what the launcher knows about the gradient reaching this level's probabilities: 0 = decide at run time,
// 1 = none (last level of the tree: the activation backward is compiled out), 2 = sigmoid level that only receives
// the uniform FiLM-pool gradient (level 0 of a two-level tree: no per-pixel term, nothing to hand to a parent).
// Kinds 1 and 2 fit 64 registers and run 512-thread CTAs (twice the warps per SM).
template <int K, int SRC, int MODE, bool BAND = false, int ACTK = 0, int THREADS = XR_THREADS>
__global__ void __launch_bounds__(THREADS, 2)
dz_rows_xreduce_kernel(const float* __restrict__ dz_hi, FusedDzArgs fa, int Wf, int H, int W, float sx, int XR_ROWS,
                       float* __restrict__ tmpx, int Hf, float sy, int band, int acc_rows, float* __restrict__ dz_lo) {
  pdl_wait();
  extern __shared__ __align__(16) float sm[];
  const int Wp = W + 4;
  float* dzs = sm;                                   // [K][XR_ROWS][Wp]
  float* wtab = sm + (size_t)K * XR_ROWS * Wp;       // [Wf][XR_MAXW]
  int* wstart = reinterpret_cast<int*>(wtab + (size_t)Wf * XR_MAXW);  // [Wf]
  int* wcnt = wstart + Wf;                           // [Wf]
  float* accs = reinterpret_cast<float*>(wcnt + Wf); // BAND: [K][acc_rows][Wf]
  const int b = blockIdx.y, tid = threadIdx.x;
  const long N = (long)H * W;
  const int band_y0 = BAND ? blockIdx.x * band : 0;
  const int band_y1 = BAND ? min(H, band_y0 + band) : H;
  const int i_lo = BAND ? make_lerp(band_y0, sy, Hf).i0 : 0;
  if constexpr (BAND)
    for (int e = tid; e < K * acc_rows * Wf; e += THREADS) accs[e] = 0.f;

  for (int j = tid; j < Wf; j += THREADS) {
    int lo, hi;
    lerp_support(j, sx, W, lo, hi);
    int first = -1, cnt = 0;
    for (int x = lo; x <= hi; ++x) {
      const float w = lerp_weight(x, sx, Wf, j);
      if (w != 0.f || first >= 0) {
        if (first < 0) first = x;
        if (cnt < XR_MAXW) wtab[j * XR_MAXW + cnt] = w;
        ++cnt;
      }
    }
    wstart[j] = first < 0 ? lo : first;
    wcnt[j] = cnt < XR_MAXW ? cnt : XR_MAXW;
  }

  // persistent over the sample's row blocks (the weight tables above are built once per CTA)
  const int vec_per_row = W / 4;
  for (int y0 = BAND ? band_y0 : blockIdx.x * XR_ROWS; y0 < band_y1; y0 += BAND ? XR_ROWS : gridDim.x * XR_ROWS) {
  const int rows = min(XR_ROWS, band_y1 - y0);
  __syncthreads();  // tables ready / previous block's readers done
  // ---- phase 1: gradient of ROWS full rows -> shared memory (4 pixels per thread and step) ----
  if constexpr (SRC == 0) {
    for (int e = tid; e < rows * vec_per_row; e += THREADS) {
      const int r = e / vec_per_row, xv = (e - r * vec_per_row) * 4;
      const size_t px = (size_t)(y0 + r) * W + xv;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(dz_hi + ((size_t)b * K + k) * N + px);
        *reinterpret_cast<float4*>(dzs + ((size_t)k * XR_ROWS + r) * Wp + xv) = v;
      }
    }
  } else {
    const LevelInfo li = load_level_info<K>(MODE == RHSEG_ACT_GROUPED ? fa.table : nullptr);
    const float gce = fa.g_ce ? __ldg(fa.g_ce) : 0.f, gdi = fa.g_dice ? __ldg(fa.g_dice) : 0.f;
    float A[K], Bc[K], Cc[K], gu[K];
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float* cf = fa.coef + ((size_t)b * K + k) * 3;
      A[k] = gce * __ldg(cf);
      Bc[k] = gdi * __ldg(cf + 1);
      Cc[k] = gdi * __ldg(cf + 2);
      gu[k] = fa.g_uniform ? (float)fa.g_uniform[b * K + k] * fa.inv_npix : 0.f;
    }
    constexpr bool NOACT = ACTK == 1, UNIF = ACTK == 2;
    static_assert(!UNIF || MODE == RHSEG_ACT_SIGMOID, "uniform-only kind: sigmoid levels");
    const bool has_act = UNIF || (!NOACT && MODE != RHSEG_ACT_ZEROS && (fa.g_uniform != nullptr || (fa.dp_pix != nullptr && fa.pix_mask != 0)));
    for (int e = tid; e < rows * vec_per_row; e += THREADS) {
      const int r = e / vec_per_row, xv = (e - r * vec_per_row) * 4;
      const size_t px = (size_t)(y0 + r) * W + xv;
      float z[K][4], t[K][4], ex[K][4], pp[K][4], o[K][4], dpar[K][4];
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const Vec<4> zv = ld_stream<4>(fa.logits + ((size_t)b * K + k) * N + px);
        const Vec<4> tv = ld_cached<4>(fa.targets + (size_t)b * fa.t_bstride + (size_t)k * fa.t_cstride + px);
        Vec<4> ev;
#pragma unroll
        for (int v = 0; v < 4; ++v) ev.v[v] = 0.f;
        if (!UNIF && has_act && fa.dp_pix && ((fa.pix_mask >> k) & 1u)) ev = ld_stream<4>(fa.dp_pix + ((size_t)b * K + k) * N + px);
#pragma unroll
        for (int v = 0; v < 4; ++v) { z[k][v] = zv.v[v]; t[k][v] = tv.v[v]; ex[k][v] = ev.v[v]; pp[k][v] = 0.f; }
      }
      if constexpr (MODE == RHSEG_ACT_GROUPED) {
        if (has_act) {
#pragma unroll
          for (int k = 0; k < K; ++k) {
            if ((li.start_mask >> k) & 1) {
              const Vec<4> pv = ld_stream<4>(fa.prev_probs + ((size_t)b * fa.K_prev + li.parent[k]) * N + px);
#pragma unroll
              for (int v = 0; v < 4; ++v) pp[k][v] = pv.v[v];
            } else {
#pragma unroll
              for (int v = 0; v < 4; ++v) pp[k][v] = pp[k > 0 ? k - 1 : 0][v];
            }
          }
        }
      }
#pragma unroll
      for (int v = 0; v < 4; ++v) {
        float zz[K], tt[K], dz[K], dP[K], ppv[K], dpv[K];
#pragma unroll
        for (int k = 0; k < K; ++k) { zz[k] = z[k][v]; tt[k] = t[k][v]; dP[k] = gu[k] + ex[k][v]; ppv[k] = pp[k][v]; dpv[k] = 0.f; }
        loss_dz_pixel<K, true>(zz, tt, A, Bc, Cc, dz);
        if (has_act) act_dz_pixel<K, MODE>(zz, dP, ppv, li.start_mask, dz, dpv);
#pragma unroll
        for (int k = 0; k < K; ++k) { o[k][v] = dz[k]; dpar[k][v] = dpv[k]; }
      }
#pragma unroll
      for (int k = 0; k < K; ++k)
        *reinterpret_cast<float4*>(dzs + ((size_t)k * XR_ROWS + r) * Wp + xv) = make_float4(o[k][0], o[k][1], o[k][2], o[k][3]);
      if constexpr (MODE == RHSEG_ACT_GROUPED) {
        if (has_act && fa.dp_prev) {
#pragma unroll
          for (int k = 0; k < K; ++k)
            if ((li.start_mask >> k) & 1) {
              float* dst = fa.dp_prev + ((size_t)b * fa.K_prev + li.parent[k]) * N + px;
              float4 cur = *reinterpret_cast<const float4*>(dst);
              cur.x += dpar[k][0]; cur.y += dpar[k][1]; cur.z += dpar[k][2]; cur.w += dpar[k][3];
              *reinterpret_cast<float4*>(dst) = cur;
            }
        }
      }
    }
  }
  __syncthreads();

  // ---- phase 2: reduce along x: tmpx[b][k][y][j] = sum_x wx(x, j) dz[k][y][x] ----
  if constexpr (BAND) {
    // thread <-> (k, j) column: it alone touches accs[k][*][j]
    for (int e = tid; e < K * Wf; e += THREADS) {
      const int k = e / Wf, j = e - k * Wf;
      const float* wt = wtab + j * XR_MAXW;
      const int n = wcnt[j], ws = wstart[j];
      for (int r = 0; r < rows; ++r) {
        const float* row = dzs + ((size_t)k * XR_ROWS + r) * Wp + ws;
        float acc = 0.f;
        for (int q = 0; q < n; ++q) acc = fmaf(wt[q], row[q], acc);
        const Lerp ly = make_lerp(y0 + r, sy, Hf);
        float* col = accs + ((long)k * acc_rows - i_lo) * Wf + j;
        col[(long)ly.i0 * Wf] = fmaf(ly.l0, acc, col[(long)ly.i0 * Wf]);
        col[(long)ly.i1 * Wf] = fmaf(ly.l1, acc, col[(long)ly.i1 * Wf]);
      }
    }
  } else
  for (int e = tid; e < K * rows * Wf; e += THREADS) {
    const int j = e % Wf;
    const int kr = e / Wf;
    const int r = kr % rows, k = kr / rows;
    const float* row = dzs + ((size_t)k * XR_ROWS + r) * Wp + wstart[j];
    const float* wt = wtab + j * XR_MAXW;
    const int n = wcnt[j];
    float acc = 0.f;
    for (int q = 0; q < n; ++q) acc = fmaf(wt[q], row[q], acc);
    tmpx[(((size_t)b * K + k) * H + y0 + r) * Wf + j] = acc;
  }
  }  // row blocks
  if constexpr (BAND) {
    __syncthreads();
    const int i_hi = make_lerp(band_y1 - 1, sy, Hf).i1;
    const int nrow = i_hi - i_lo + 1;
    for (int e = tid; e < K * nrow * Wf; e += THREADS) {
      const int j = e % Wf;
      const int kr = e / Wf;
      const int ri = kr % nrow, k = kr / nrow;
      atomicAdd(dz_lo + (((size_t)b * K + k) * Hf + i_lo + ri) * Wf + j, accs[((size_t)k * acc_rows + ri) * Wf + j]);
    }
  }
}

// pass 2: dz_lo[b][k][i][j] = sum_y wy(y, i) tmpx[b][k][y][j]
__global__ void __launch_bounds__(256)
yreduce_kernel(const float* __restrict__ tmpx, int Hf, int Wf, int H, float sy, long total, float* __restrict__ dz_lo) {
  pdl_wait();
  const long idx = (long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int j = (int)(idx % Wf);
  const int i = (int)((idx / Wf) % Hf);
  const long bk = idx / ((long)Wf * Hf);
  int lo, hi;
  lerp_support(i, sy, H, lo, hi);
  const float* col = tmpx + (size_t)bk * H * Wf + j;
  float acc = 0.f;
  for (int y = lo; y <= hi; ++y) {
    const float w = lerp_weight(y, sy, Hf, i);
    if (w != 0.f) acc = fmaf(w, __ldg(col + (size_t)y * Wf), acc);
  }
  dz_lo[idx] = acc;
}

// Full-resolution donors (UNet): the same fused per-pixel gradient, written out once as dz.
template <int K, int VEC, int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS)
dz_fullres_fused_kernel(FusedDzArgs fa, long N, float* __restrict__ dz_out) {
  pdl_wait();
  const int b = blockIdx.y;
  const long px = ((long)blockIdx.x * THREADS + threadIdx.x) * VEC;
  if (px >= N) return;
  const LevelInfo li = load_level_info<K>(MODE == RHSEG_ACT_GROUPED ? fa.table : nullptr);
  const float gce = fa.g_ce ? __ldg(fa.g_ce) : 0.f, gdi = fa.g_dice ? __ldg(fa.g_dice) : 0.f;
  float A[K], Bc[K], Cc[K], gu[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const float* cf = fa.coef + ((size_t)b * K + k) * 3;
    A[k] = gce * __ldg(cf);
    Bc[k] = gdi * __ldg(cf + 1);
    Cc[k] = gdi * __ldg(cf + 2);
    gu[k] = fa.g_uniform ? (float)fa.g_uniform[b * K + k] * fa.inv_npix : 0.f;
  }
  const bool has_act = MODE != RHSEG_ACT_ZEROS && (fa.g_uniform != nullptr || (fa.dp_pix != nullptr && fa.pix_mask != 0));
  float z[K][VEC], t[K][VEC], e[K][VEC], pp[K][VEC], o[K][VEC], dpar[K][VEC];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const Vec<VEC> zv = ld_stream<VEC>(fa.logits + ((size_t)b * K + k) * N + px);
    const Vec<VEC> tv = ld_cached<VEC>(fa.targets + (size_t)b * fa.t_bstride + (size_t)k * fa.t_cstride + px);
    Vec<VEC> ev;
#pragma unroll
    for (int v = 0; v < VEC; ++v) ev.v[v] = 0.f;
    if (has_act && fa.dp_pix && ((fa.pix_mask >> k) & 1u)) ev = ld_stream<VEC>(fa.dp_pix + ((size_t)b * K + k) * N + px);
#pragma unroll
    for (int v = 0; v < VEC; ++v) { z[k][v] = zv.v[v]; t[k][v] = tv.v[v]; e[k][v] = ev.v[v]; pp[k][v] = 0.f; }
  }
  if constexpr (MODE == RHSEG_ACT_GROUPED) {
    if (has_act) {
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if ((li.start_mask >> k) & 1) {
          const Vec<VEC> pv = ld_stream<VEC>(fa.prev_probs + ((size_t)b * fa.K_prev + li.parent[k]) * N + px);
#pragma unroll
          for (int v = 0; v < VEC; ++v) pp[k][v] = pv.v[v];
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) pp[k][v] = pp[k > 0 ? k - 1 : 0][v];
        }
      }
    }
  }
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    float zz[K], tt[K], dz[K], dP[K], ppv[K], dpv[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { zz[k] = z[k][v]; tt[k] = t[k][v]; dP[k] = gu[k] + e[k][v]; ppv[k] = pp[k][v]; dpv[k] = 0.f; }
    loss_dz_pixel<K, true>(zz, tt, A, Bc, Cc, dz);
    if (has_act) act_dz_pixel<K, MODE>(zz, dP, ppv, li.start_mask, dz, dpv);
#pragma unroll
    for (int k = 0; k < K; ++k) { o[k][v] = dz[k]; dpar[k][v] = dpv[k]; }
  }
#pragma unroll
  for (int k = 0; k < K; ++k) {
    Vec<VEC> r;
#pragma unroll
    for (int v = 0; v < VEC; ++v) r.v[v] = o[k][v];
    *reinterpret_cast<Vec<VEC>*>(dz_out + ((size_t)b * K + k) * N + px) = r;  // re-read by the conv backward: keep cached
  }
  if constexpr (MODE == RHSEG_ACT_GROUPED) {
    if (has_act && fa.dp_prev) {
#pragma unroll
      for (int k = 0; k < K; ++k)
        if ((li.start_mask >> k) & 1) {
          float* dst = fa.dp_prev + ((size_t)b * fa.K_prev + li.parent[k]) * N + px;
          Vec<VEC> cur = *reinterpret_cast<const Vec<VEC>*>(dst);
#pragma unroll
          for (int v = 0; v < VEC; ++v) cur.v[v] += dpar[k][v];
          *reinterpret_cast<Vec<VEC>*>(dst) = cur;
        }
    }
  }
}

// host: worst-case support extent of a tile of `t` inputs (mirrors lerp_support)
static int region_extent(int t, float scale, int out_size) {
  if (scale <= 0.f) return out_size;
  const int e = (int)ceilf((float)(t + 1) / scale) + 5;
  return e < out_size ? e : out_size;
}

template <int K, int SRC, int MODE>
static int launch_adjoint(const float* dz_hi, const FusedDzArgs& fa, int B, int Hf, int Wf, int H, int W, float* dz_lo,
                          float* tmpx, bool prezeroed, cudaStream_t st) {
  const float sy = H > 1 ? (float)(Hf - 1) / (float)(H - 1) : 0.f;
  const float sx = W > 1 ? (float)(Wf - 1) / (float)(W - 1) : 0.f;
  {
    auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    bool rows_ok = (tmpx != nullptr || prezeroed) && W % 4 == 0 && sx > 0.f && 2.0f / sx + 3.0f <= (float)XR_MAXW;
    if (SRC == 0) rows_ok = rows_ok && al(dz_hi);
    else rows_ok = rows_ok && al(fa.logits) && al(fa.targets) && fa.t_bstride % 4 == 0 && fa.t_cstride % 4 == 0 &&
                   al(fa.prev_probs) && al(fa.dp_pix) && al(fa.dp_prev);
    constexpr int ROWS = 2;  // rows per block: small blocks, the persistent loop balances them over one wave
    const size_t smem = ((size_t)K * ROWS * (W + 4) + (size_t)Wf * XR_MAXW + 2 * (size_t)Wf) * sizeof(float);
    const bool no_band = getenv("RHSEG_NO_BAND_ADJOINT") != nullptr;
    if (rows_ok && prezeroed && !no_band && sy > 0.f && sy <= 1.0f) {
      // band kernel: x- and y-reduction in one pass into the pre-zeroed dz_lo
      // without the activation backward the kernel needs 64 registers: 512-thread CTAs double the warps per SM
      const bool no_pix = fa.dp_pix == nullptr || fa.pix_mask == 0;
      const bool no_act = SRC == 1 && (MODE == RHSEG_ACT_ZEROS || (fa.g_uniform == nullptr && no_pix));
      const bool unif = SRC == 1 && MODE == RHSEG_ACT_SIGMOID && fa.g_uniform != nullptr && no_pix;
      const int nthreads = (no_act || unif) ? 2 * XR_THREADS : XR_THREADS;
      constexpr int T2 = (SRC == 1) ? 2 * XR_THREADS : XR_THREADS;
      auto kern = no_act ? dz_rows_xreduce_kernel<K, SRC, MODE, true, (SRC == 1 ? 1 : 0), T2>
                : unif   ? dz_rows_xreduce_kernel<K, SRC, MODE, true, ((SRC == 1 && MODE == RHSEG_ACT_SIGMOID) ? 2 : 0),
                                                  ((SRC == 1 && MODE == RHSEG_ACT_SIGMOID) ? T2 : XR_THREADS)>
                         : dz_rows_xreduce_kernel<K, SRC, MODE, true, 0, XR_THREADS>;
      // >= taps per low-res row (ceil(2/sy) - 1 would do): at most two CTAs feed one low-res row
      const int min_band = std::max(2, (int)ceilf(2.0f / sy));
      // rows per phase-1 pass: the one that wastes the fewest threads (a pass handles rows * W/4 four-pixel items
      // with nthreads threads: 2 rows of 620 px keep only 60 % of them busy, 3 rows 91 %)
      const int vpr = W / 4;
      int sub = 1;
      double best_util = 0.0;
      for (int r = 1; r <= 4; ++r) {
        const int items = r * vpr;
        const double util = (double)items / (double)(((items + nthreads - 1) / nthreads) * nthreads);
        if (util > best_util + 1e-9) { best_util = util; sub = r; }
      }
      auto smem_for = [&](int band) {
        const int acc_rows = (int)floorf(sy * (float)band) + 4;
        return ((size_t)K * sub * (W + 4) + (size_t)Wf * XR_MAXW + 2 * (size_t)Wf + (size_t)K * acc_rows * Wf) * sizeof(float);
      };
      auto round_band = [&](int band) { return ((std::max(band, min_band) + sub - 1) / sub) * sub; };
      int per_sm = 0;
      const size_t smem0 = smem_for(round_band(min_band));
      if (smem0 <= 200 * 1024) {
        if (smem0 > 48 * 1024) RHSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem0));
        RHSEG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, nthreads, smem0));
      }
      if (per_sm >= 1) {
        const long slots = std::max<long>(1, (long)per_sm * device_sm_count() / B);
        const int band = round_band((int)((H + slots - 1) / slots));
        const int acc_rows = (int)floorf(sy * (float)band) + 4;
        const size_t smem_b = smem_for(band);
        if (smem_b <= 200 * 1024) {
          if (smem_b > 48 * 1024) RHSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_b));
          dim3 grid((unsigned)((H + band - 1) / band), B);
          launch_pdl(kern, dim3(grid), dim3(nthreads), smem_b, st, dz_hi, fa, Wf, H, W, sx, sub, (float*)nullptr, Hf, sy, band,
                     acc_rows, dz_lo);
          RHSEG_LAUNCH_CHECK();
          return RHSEG_OK;
        }
      }
    }
    if (prezeroed && tmpx == nullptr) rows_ok = false;  // two-kernel row path needs the workspace
    auto kern = dz_rows_xreduce_kernel<K, SRC, MODE>;
    if (rows_ok && smem <= 200 * 1024) {
      if (smem > 48 * 1024) RHSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      int per_sm = 0;
      RHSEG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, XR_THREADS, smem));
      if (per_sm < 1) per_sm = 1;
      const long slots = std::max<long>(1, (long)per_sm * device_sm_count() / B);
      dim3 grid(balanced_grid((H + ROWS - 1) / ROWS, slots), B);
      launch_pdl(kern, dim3(grid), dim3(XR_THREADS), smem, st, dz_hi, fa, Wf, H, W, sx, ROWS, tmpx, Hf, sy, 0, 0, (float*)nullptr);
      RHSEG_LAUNCH_CHECK();
      const long total = (long)B * K * Hf * Wf;
      launch_pdl(yreduce_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, tmpx, Hf, Wf, H, sy, total, dz_lo);
      RHSEG_LAUNCH_CHECK();
      return RHSEG_OK;
    }
  }
  const int ry_max = region_extent(ADJ_TH, sy, H), rx_max = region_extent(ADJ_TW, sx, W) | 1;  // odd pitch: fewer bank conflicts
  const size_t smem = ((size_t)K * ry_max * rx_max + (size_t)K * ry_max * ADJ_TW) * sizeof(float);
  if (smem > 200 * 1024) return RHSEG_ERR_UNSUPPORTED;  // upsampling factor too large for the tiled kernel
  if (sy > 0.f && 2.0f / sy + 4.0f > (float)ADJ_MAXW) return RHSEG_ERR_UNSUPPORTED;
  if (sx > 0.f && 2.0f / sx + 4.0f > (float)ADJ_MAXW) return RHSEG_ERR_UNSUPPORTED;
  if ((sy <= 0.f && H > ADJ_MAXW) || (sx <= 0.f && W > ADJ_MAXW)) return RHSEG_ERR_UNSUPPORTED;
  auto kern = upsample_adjoint_tiled_kernel<K, SRC, MODE>;
  if (smem > 48 * 1024) RHSEG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((Wf + ADJ_TW - 1) / ADJ_TW, (Hf + ADJ_TH - 1) / ADJ_TH, B);
  launch_pdl(kern, dim3(grid), dim3(ADJ_THREADS), smem, st, dz_hi, fa, Hf, Wf, H, W, sy, sx, ry_max, rx_max, dz_lo);
  RHSEG_LAUNCH_CHECK();
  return RHSEG_OK;
}

}  // namespace rhseg

using namespace rhseg;

extern "C" int rhseg_upsample_adjoint(const float* dz_hi, int B, int K, int Hf, int Wf, int H, int W, float* dz_lo,
                                      float* tmp, int flags, void* stream) {
  if (!dz_hi || !dz_lo || B <= 0 || Hf <= 0 || Wf <= 0 || H <= 0 || W <= 0) return RHSEG_ERR_ARG;
  if (K < 1 || K > RHSEG_KERNEL_MAX_K) return RHSEG_ERR_UNSUPPORTED;
  FusedDzArgs fa{};
  RHSEG_DISPATCH_K(K, return (launch_adjoint<KK, 0, 0>(dz_hi, fa, B, Hf, Wf, H, W, dz_lo, tmp, (flags & RHSEG_DZ_PREZEROED) != 0, (cudaStream_t)stream)));
  return RHSEG_OK;
}

extern "C" int rhseg_head_dz_lowres_fused(const float* logits, const float* targets, long t_bstride, long t_cstride,
                                          const float* coef, const float* g_ce, const float* g_dice,
                                          const float* prev_probs, const int32_t* table, const double* g_uniform,
                                          double inv_npix, const float* dp_pix, uint32_t pix_mask, int B, int K,
                                          int K_prev, int Hf, int Wf, int H, int W, int act_mode, float* dz_lo,
                                          float* dp_prev, float* tmp, int flags, void* stream) {
  if (!logits || !targets || !coef || !dz_lo || B <= 0 || Hf <= 0 || Wf <= 0 || H <= 0 || W <= 0) return RHSEG_ERR_ARG;
  if (K < 1 || K > RHSEG_KERNEL_MAX_K) return RHSEG_ERR_UNSUPPORTED;
  if (act_mode == RHSEG_ACT_GROUPED && (!prev_probs || !table)) return RHSEG_ERR_ARG;
  FusedDzArgs fa{logits, targets, t_bstride, t_cstride, coef, g_ce, g_dice, prev_probs, table, g_uniform,
                 (float)inv_npix, dp_pix, pix_mask, dp_prev, K_prev};
  cudaStream_t st = (cudaStream_t)stream;
  const bool pz = (flags & RHSEG_DZ_PREZEROED) != 0;
  RHSEG_DISPATCH_K(K, {
    if (act_mode == RHSEG_ACT_SIGMOID) return launch_adjoint<KK, 1, RHSEG_ACT_SIGMOID>(nullptr, fa, B, Hf, Wf, H, W, dz_lo, tmp, pz, st);
    if (act_mode == RHSEG_ACT_GROUPED) return launch_adjoint<KK, 1, RHSEG_ACT_GROUPED>(nullptr, fa, B, Hf, Wf, H, W, dz_lo, tmp, pz, st);
    return launch_adjoint<KK, 1, RHSEG_ACT_ZEROS>(nullptr, fa, B, Hf, Wf, H, W, dz_lo, tmp, pz, st);
  });
  return RHSEG_OK;
}

extern "C" int rhseg_head_dz_fullres_fused(const float* logits, const float* targets, long t_bstride, long t_cstride,
                                           const float* coef, const float* g_ce, const float* g_dice,
                                           const float* prev_probs, const int32_t* table, const double* g_uniform,
                                           double inv_npix, const float* dp_pix, uint32_t pix_mask, int B, int K,
                                           int K_prev, int n_pix, int act_mode, float* dz_out, float* dp_prev,
                                           void* stream) {
  if (!logits || !targets || !coef || !dz_out || B <= 0 || n_pix <= 0) return RHSEG_ERR_ARG;
  if (K < 1 || K > RHSEG_KERNEL_MAX_K) return RHSEG_ERR_UNSUPPORTED;
  if (act_mode == RHSEG_ACT_GROUPED && (!prev_probs || !table)) return RHSEG_ERR_ARG;
  FusedDzArgs fa{logits, targets, t_bstride, t_cstride, coef, g_ce, g_dice, prev_probs, table, g_uniform,
                 (float)inv_npix, dp_pix, pix_mask, dp_prev, K_prev};
  cudaStream_t st = (cudaStream_t)stream;
  const long N = n_pix;
  constexpr int THREADS = 256;
  auto al = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  const bool v4 = (N % 4 == 0) && al(logits) && al(targets) && t_bstride % 4 == 0 && t_cstride % 4 == 0 && al(dz_out) &&
                  al(prev_probs) && al(dp_pix) && al(dp_prev);
  RHSEG_DISPATCH_K(K, {
    if (v4) {
      dim3 grid((unsigned)((N / 4 + THREADS - 1) / THREADS), B);
      if (act_mode == RHSEG_ACT_SIGMOID) launch_pdl(dz_fullres_fused_kernel<KK, 4, RHSEG_ACT_SIGMOID, THREADS>, dim3(grid), dim3(THREADS), 0, st, fa, N, dz_out);
      else if (act_mode == RHSEG_ACT_GROUPED) launch_pdl(dz_fullres_fused_kernel<KK, 4, RHSEG_ACT_GROUPED, THREADS>, dim3(grid), dim3(THREADS), 0, st, fa, N, dz_out);
      else launch_pdl(dz_fullres_fused_kernel<KK, 4, RHSEG_ACT_ZEROS, THREADS>, dim3(grid), dim3(THREADS), 0, st, fa, N, dz_out);
    } else {
      dim3 grid((unsigned)((N + THREADS - 1) / THREADS), B);
      if (act_mode == RHSEG_ACT_SIGMOID) launch_pdl(dz_fullres_fused_kernel<KK, 1, RHSEG_ACT_SIGMOID, THREADS>, dim3(grid), dim3(THREADS), 0, st, fa, N, dz_out);
      else if (act_mode == RHSEG_ACT_GROUPED) launch_pdl(dz_fullres_fused_kernel<KK, 1, RHSEG_ACT_GROUPED, THREADS>, dim3(grid), dim3(THREADS), 0, st, fa, N, dz_out);
      else launch_pdl(dz_fullres_fused_kernel<KK, 1, RHSEG_ACT_ZEROS, THREADS>, dim3(grid), dim3(THREADS), 0, st, fa, N, dz_out);
    }
  });
  RHSEG_LAUNCH_CHECK();
  return RHSEG_OK;
}

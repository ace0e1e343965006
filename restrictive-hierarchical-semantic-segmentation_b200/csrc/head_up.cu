// HRNet hi-res pass: bilinear (align_corners=True) upsample of the low-res logits fused with the activation,
// the composition, the FiLM pool sums and (optionally) the level's training evaluation.
// Reference semantics: Models/models.py:766/:776 (upsample), :767/:779-796 (activation, composition).
#include <algorithm>
#include <cstdlib>
#include "common.cuh"
#include "eval_accum.cuh"
#include "head_common.cuh"

namespace rhseg {

// ------------------------------------------------------------------------------------
// HRNet hi-res pass: bilinear (align_corners=True) upsample of the low-res logits fused with
// the activation.  Index/lambda arithmetic follows ATen's upsample_bilinear2d (fp32 scale =
// (in-1)/(out-1), src = scale*dst, i0 = (int)src, lambda1 = src - i0).
// ------------------------------------------------------------------------------------
template <int K, int VEC, int MODE, int THREADS>
__global__ void __launch_bounds__(THREADS)
upsample_act_kernel(const float* __restrict__ z_lo, const float* __restrict__ prev_probs,
                    const int32_t* __restrict__ table, int Hf, int Wf, int H, int W, int K_prev,
                    float sy, float sx, long total_vec, float* __restrict__ logits,
                    float* __restrict__ probs, double* __restrict__ psum, int vec_per_sample) {
  pdl_wait();
  __shared__ float red[(THREADS / 32) * K];
  const int b = blockIdx.y;
  const long vi = (long)blockIdx.x * THREADS + threadIdx.x;  // vector index inside the sample
  const bool ok = vi < vec_per_sample;
  const int wv = W / VEC;
  const int y = ok ? (int)(vi / wv) : 0;
  const int x0 = ok ? (int)(vi - (long)y * wv) * VEC : 0;
  const long N = (long)H * W, Nf = (long)Hf * Wf;
  const long px = (long)y * W + x0;

  float z[K][VEC], pp[K][VEC], prob[K][VEC];
  const LevelInfo li = load_level_info<K>(MODE == RHSEG_ACT_GROUPED ? table : nullptr);
  if (ok) {
    const Lerp ly = make_lerp(y, sy, Hf);
    Lerp lx[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) lx[v] = make_lerp(x0 + v, sx, Wf);
    const float* zb = z_lo + (size_t)b * K * Nf;
    // neighbour offsets are shared by all K channels
    int o00[VEC], o01[VEC], o10[VEC], o11[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
      o00[v] = ly.i0 * Wf + lx[v].i0; o01[v] = ly.i0 * Wf + lx[v].i1;
      o10[v] = ly.i1 * Wf + lx[v].i0; o11[v] = ly.i1 * Wf + lx[v].i1;
    }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      const float* zk = zb + (size_t)k * Nf;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const float a = __ldg(zk + o00[v]), bq = __ldg(zk + o01[v]);
        const float c = __ldg(zk + o10[v]), d = __ldg(zk + o11[v]);
        z[k][v] = ly.l0 * (lx[v].l0 * a + lx[v].l1 * bq) + ly.l1 * (lx[v].l0 * c + lx[v].l1 * d);
      }
    }
    if constexpr (MODE == RHSEG_ACT_GROUPED) {
      const float* pb = prev_probs + (size_t)b * K_prev * N;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if ((li.start_mask >> k) & 1) {
          const Vec<VEC> t = ld_cached<VEC>(pb + (size_t)li.parent[k] * N + px);
#pragma unroll
          for (int v = 0; v < VEC; ++v) pp[k][v] = t.v[v];
        } else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) pp[k][v] = pp[k > 0 ? k - 1 : 0][v];
        }
      }
    }
  } else {
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int v = 0; v < VEC; ++v) { z[k][v] = 0.f; pp[k][v] = 0.f; }
  }
  activate<K, VEC, MODE>(z, pp, li.start_mask, prob);
  float ps[K];
#pragma unroll
  for (int k = 0; k < K; ++k) {
    ps[k] = 0.f;
    if (ok) {
      Vec<VEC> zo, po;
#pragma unroll
      for (int v = 0; v < VEC; ++v) { zo.v[v] = z[k][v]; po.v[v] = prob[k][v]; ps[k] += po.v[v]; }
      *reinterpret_cast<Vec<VEC>*>(logits + ((size_t)b * K + k) * N + px) = zo;
      *reinterpret_cast<Vec<VEC>*>(probs + ((size_t)b * K + k) * N + px) = po;
    }
  }
  block_psum<K, THREADS / 32>(ps, psum + (size_t)b * K, red, [] { __syncthreads(); });
}

// Tiled variant for upsampling factors >= 1 (the HRNet case): a CTA of 16x16 threads owns a
// 16 x (16*VEC) hi-res tile, stages the low-res logit patch it reads in shared memory and
// interpolates from there (shared-memory loads with immediate offsets instead of 64 global loads
// with 64-bit address arithmetic per thread).
// EVAL: additionally run the per-level training evaluation (eval_accum.cuh) on the pixels while their
// logits are in registers: statistics, prediction index map, confusion matrix, consistency sums.

// EVALK: 0 = no evaluation, 1 = evaluate (no consistency inputs), 2 = evaluate with the consistency sums; together
// with MODE this fixes the level kind at compile time (EvalAccum's CT).
template <int K, int VEC, int MODE, int EVALK>
__global__ void __launch_bounds__(256, VEC == 2 ? 3 : 2)
upsample_act_tiled_kernel(const float* __restrict__ z_lo, const float* __restrict__ prev_probs,
                          const int32_t* __restrict__ table, int Hf, int Wf, int H, int W, int K_prev, float sy,
                          float sx, int tiles_x, int tiles_per_sample, float* __restrict__ logits,
                          float* __restrict__ probs, double* __restrict__ psum, EvalArgs ea) {
  pdl_wait();
  constexpr bool EVAL = EVALK != 0;
  constexpr int CT = MODE == RHSEG_ACT_SIGMOID ? 0 : (EVALK == 2 ? 2 : 1);
  constexpr int TH = 16, TW = 16 * VEC;
  constexpr int PH = TH + 2, PW = TW + 2;  // patch bound for scale <= 1
  // double buffered when it fits the static limit: the next tile's patch streams in (cp.async) while this one is used
  constexpr int NBUF = (2 * K * PH * PW * 4 <= 40 * 1024) ? 2 : 1;
  __shared__ float patch2[NBUF][K][PH][PW];
  __shared__ float red[8 * K];
  __shared__ float ered[EVAL ? 8 * EvalAccum<K>::NACC : 1];
  __shared__ int hist[EVAL ? (K + 1) * (K + 1) : 1];
  const int b = blockIdx.y, tid = threadIdx.x;
  EvalAccum<K, CT> ev;
  if constexpr (EVAL)
    ev.init(ea.child, ea.child && ea.prev_idx != nullptr && ea.parent_targets != nullptr, table, hist, 256);
  const LevelInfo li = load_level_info<K>(MODE == RHSEG_ACT_GROUPED ? table : nullptr);
  const long N = (long)H * W, Nf = (long)Hf * Wf;
  const float* zb = z_lo + (size_t)b * K * Nf;
  float ps[K];
#pragma unroll
  for (int k = 0; k < K; ++k) ps[k] = 0.f;

  // low-res support of one tile -> shared memory, asynchronously (4-byte cp.async: HRNet planes are only 4-byte aligned)
  auto issue_patch = [&](int tile, int buf) {
    const int ty0 = (tile / tiles_x) * TH, tx0 = (tile % tiles_x) * TW;
    const int ylast = min(ty0 + TH, H) - 1, xlast = min(tx0 + TW, W) - 1;
    const int r0 = (int)(sy * (float)ty0), c0 = (int)(sx * (float)tx0);
    const int r1 = min((int)(sy * (float)ylast) + 1, Hf - 1), c1 = min((int)(sx * (float)xlast) + 1, Wf - 1);
    const int ph = r1 - r0 + 1, pw = c1 - c0 + 1;
    const int lane = tid & 31, warp = tid >> 5;
    for (int rr = warp; rr < ph; rr += 8)
      for (int cc = lane; cc < pw; cc += 32) {
        const float* src = zb + (size_t)(r0 + rr) * Wf + c0 + cc;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          const uint32_t dst = (uint32_t)__cvta_generic_to_shared(&patch2[buf][k][rr][cc]);
          asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src + (size_t)k * Nf) : "memory");
        }
      }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // persistent over the sample's tiles: pool sums / evaluation statistics are reduced once per CTA
  int buf = 0;
  if (NBUF == 2 && (int)blockIdx.x < tiles_per_sample) issue_patch(blockIdx.x, 0);
  for (int tile = blockIdx.x; tile < tiles_per_sample; tile += gridDim.x, buf ^= (NBUF - 1)) {
    const int ty0 = (tile / tiles_x) * TH, tx0 = (tile % tiles_x) * TW;
    const int r0 = (int)(sy * (float)ty0), c0 = (int)(sx * (float)tx0);
    if constexpr (NBUF == 1) {
      __syncthreads();  // previous tile's readers are done with the patch
      issue_patch(tile, 0);
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();  // this tile's patch has landed; everybody is done reading the other buffer
    if (NBUF == 2 && tile + (int)gridDim.x < tiles_per_sample) issue_patch(tile + gridDim.x, buf ^ 1);
    float (*patch)[PH][PW] = patch2[buf];
    const int y = ty0 + (tid >> 4), x0 = tx0 + (tid & 15) * VEC;
    const bool ok = y < H && x0 < W;  // W % VEC == 0 guaranteed by the launcher
    const long px = (long)y * W + x0;
    float z[K][VEC], pp[K][VEC], prob[K][VEC];
    if (ok) {
      const Lerp ly = make_lerp(y, sy, Hf);
      const int row0 = (ly.i0 - r0) * PW, row1 = (ly.i1 - r0) * PW;
      const float* pbase = &patch[0][0][0];
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const Lerp lx = make_lerp(x0 + v, sx, Wf);
        const float* p00 = pbase + row0 + (lx.i0 - c0);
        const float* p01 = pbase + row0 + (lx.i1 - c0);
        const float* p10 = pbase + row1 + (lx.i0 - c0);
        const float* p11 = pbase + row1 + (lx.i1 - c0);
#pragma unroll
        for (int k = 0; k < K; ++k)
          z[k][v] = ly.l0 * (lx.l0 * p00[k * PH * PW] + lx.l1 * p01[k * PH * PW]) +
                    ly.l1 * (lx.l0 * p10[k * PH * PW] + lx.l1 * p11[k * PH * PW]);
      }
      if constexpr (MODE == RHSEG_ACT_GROUPED) {
        const float* pb = prev_probs + (size_t)b * K_prev * N;
#pragma unroll
        for (int k = 0; k < K; ++k) {
          if ((li.start_mask >> k) & 1) {
            const Vec<VEC> t = ld_cached<VEC>(pb + (size_t)li.parent[k] * N + px);
#pragma unroll
            for (int v = 0; v < VEC; ++v) pp[k][v] = t.v[v];
          } else {
#pragma unroll
            for (int v = 0; v < VEC; ++v) pp[k][v] = pp[k > 0 ? k - 1 : 0][v];
          }
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < K; ++k)
#pragma unroll
        for (int v = 0; v < VEC; ++v) { z[k][v] = 0.f; pp[k][v] = 0.f; }
    }
    activate<K, VEC, MODE>(z, pp, li.start_mask, prob);
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (ok) {
        Vec<VEC> zo, po;
#pragma unroll
        for (int v = 0; v < VEC; ++v) { zo.v[v] = z[k][v]; po.v[v] = prob[k][v]; ps[k] += po.v[v]; }
        *reinterpret_cast<Vec<VEC>*>(logits + ((size_t)b * K + k) * N + px) = zo;
        *reinterpret_cast<Vec<VEC>*>(probs + ((size_t)b * K + k) * N + px) = po;
      }
    }
    if constexpr (EVAL) {
      float t[K][VEC], ptv[K][VEC];
      unsigned char pidx[VEC], my_idx[VEC];
      ev.template load_targets<VEC>(ea.targets, ea.t_bstride, ea.t_cstride, ea.parent_targets, ea.pt_bstride, ea.pt_cstride,
                                    ea.prev_idx, b, N, px, ok, t, ptv, pidx);
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        float zz[K], tt[K], pt[K];
#pragma unroll
        for (int k = 0; k < K; ++k) { zz[k] = z[k][v]; tt[k] = t[k][v]; pt[k] = ptv[k][v]; }
        my_idx[v] = (unsigned char)ev.pixel(zz, tt, pt, pidx[v], ok);
      }
      ev.template store_idx<VEC>(ea.idx_out, b, N, px, ok, my_idx);
    }
  }
  block_psum<K, 8>(ps, psum + (size_t)b * K, red, [] { __syncthreads(); });
  if constexpr (EVAL) ev.template finish<8>(ered, hist, ea.stats, ea.cons, ea.conf, table, b);
}


template <int K, int MODE>
static int fwd_upsampled(const float* z_lo, const float* prev_probs, const int32_t* table, int B, int Hf, int Wf,
                         int H, int W, int K_prev, float* logits, float* probs, double* psum, cudaStream_t st,
                         const EvalArgs* ea = nullptr) {
  const float sy = H > 1 ? (float)(Hf - 1) / (float)(H - 1) : 0.f;
  const float sx = W > 1 ? (float)(Wf - 1) / (float)(W - 1) : 0.f;
  constexpr int THREADS = 256;
  if (sy <= 1.0f && sx <= 1.0f) {  // upsampling: tiled kernel
    auto al4 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3u) == 0; };
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    bool v4 = W % 4 == 0;
    if (ea) v4 = v4 && al16(ea->targets) && ea->t_bstride % 4 == 0 && ea->t_cstride % 4 == 0 && al4(ea->prev_idx) && al4(ea->idx_out) &&
                 (!ea->parent_targets || (al16(ea->parent_targets) && ea->pt_bstride % 4 == 0 && ea->pt_cstride % 4 == 0));
    const EvalArgs none{};
    static int tune = -1;
    if (tune < 0) { const char* e = getenv("RHSEG_TUNE_UPACT"); tune = e ? atoi(e) : 0; }
    if (tune == 1 && v4) {  // 2 pixels per thread: fewer live registers, more resident warps
      const int slots2 = std::max(1, device_sm_count() * 3 / B);
      const int tiles_x = (W + 31) / 32, tiles = tiles_x * ((H + 15) / 16);
      dim3 grid(balanced_grid(tiles, slots2), B);
      if (ea) {
        if (ea->prev_idx && ea->parent_targets && ea->child) launch_pdl(upsample_act_tiled_kernel<K, 2, MODE, 2>, dim3(grid), dim3(256), 0, st, z_lo, prev_probs, table, Hf, Wf, H, W, K_prev, sy, sx, tiles_x, tiles, logits, probs, psum, *ea);
        else launch_pdl(upsample_act_tiled_kernel<K, 2, MODE, 1>, dim3(grid), dim3(256), 0, st, z_lo, prev_probs, table, Hf, Wf, H, W, K_prev, sy, sx, tiles_x, tiles, logits, probs, psum, *ea);
      }
      else launch_pdl(upsample_act_tiled_kernel<K, 2, MODE, 0>, dim3(grid), dim3(256), 0, st, z_lo, prev_probs, table, Hf, Wf, H, W, K_prev, sy, sx, tiles_x, tiles, logits, probs, psum, none);
      RHSEG_LAUNCH_CHECK();
      return RHSEG_OK;
    }
    const int slots = std::max(1, device_sm_count() * 2 / B);  // CTAs per sample: one resident wave at 2 CTAs/SM
    if (v4) {
      const int tiles_x = (W + 63) / 64, tiles = tiles_x * ((H + 15) / 16);
      dim3 grid(balanced_grid(tiles, slots), B);
      if (ea) {
        if (ea->prev_idx && ea->parent_targets && ea->child) launch_pdl(upsample_act_tiled_kernel<K, 4, MODE, 2>, dim3(grid), dim3(256), 0, st, z_lo, prev_probs, table, Hf, Wf, H, W, K_prev, sy, sx, tiles_x, tiles, logits, probs, psum, *ea);
        else launch_pdl(upsample_act_tiled_kernel<K, 4, MODE, 1>, dim3(grid), dim3(256), 0, st, z_lo, prev_probs, table, Hf, Wf, H, W, K_prev, sy, sx, tiles_x, tiles, logits, probs, psum, *ea);
      }
      else launch_pdl(upsample_act_tiled_kernel<K, 4, MODE, 0>, dim3(grid), dim3(256), 0, st, z_lo, prev_probs, table, Hf, Wf, H, W, K_prev, sy, sx, tiles_x, tiles, logits, probs, psum, none);
    } else {
      const int tiles_x = (W + 15) / 16, tiles = tiles_x * ((H + 15) / 16);
      dim3 grid(balanced_grid(tiles, slots), B);
      if (ea) {
        if (ea->prev_idx && ea->parent_targets && ea->child) launch_pdl(upsample_act_tiled_kernel<K, 1, MODE, 2>, dim3(grid), dim3(256), 0, st, z_lo, prev_probs, table, Hf, Wf, H, W, K_prev, sy, sx, tiles_x, tiles, logits, probs, psum, *ea);
        else launch_pdl(upsample_act_tiled_kernel<K, 1, MODE, 1>, dim3(grid), dim3(256), 0, st, z_lo, prev_probs, table, Hf, Wf, H, W, K_prev, sy, sx, tiles_x, tiles, logits, probs, psum, *ea);
      }
      else launch_pdl(upsample_act_tiled_kernel<K, 1, MODE, 0>, dim3(grid), dim3(256), 0, st, z_lo, prev_probs, table, Hf, Wf, H, W, K_prev, sy, sx, tiles_x, tiles, logits, probs, psum, none);
    }
    RHSEG_LAUNCH_CHECK();
    return RHSEG_OK;
  }
  if (ea) return RHSEG_ERR_UNSUPPORTED;  // fused evaluation exists for upsampling heads only
  if (W % 4 == 0) {
    const int vps = H * (W / 4);
    dim3 grid((vps + THREADS - 1) / THREADS, B);
    launch_pdl(upsample_act_kernel<K, 4, MODE, THREADS>, dim3(grid), dim3(THREADS), 0, st, z_lo, prev_probs, table, Hf, Wf, H, W, K_prev,
                                                                       sy, sx, 0, logits, probs, psum, vps);
  } else {
    const int vps = H * W;
    dim3 grid((vps + THREADS - 1) / THREADS, B);
    launch_pdl(upsample_act_kernel<K, 1, MODE, THREADS>, dim3(grid), dim3(THREADS), 0, st, z_lo, prev_probs, table, Hf, Wf, H, W, K_prev,
                                                                       sy, sx, 0, logits, probs, psum, vps);
  }
  RHSEG_LAUNCH_CHECK();
  return RHSEG_OK;
}


int fwd_upsampled_dispatch(int K, int act_mode, const float* z_lo, const float* prev_probs, const int32_t* table, int B,
                           int Hf, int Wf, int H, int W, int K_prev, float* logits, float* probs, double* psum,
                           cudaStream_t st, const EvalArgs* ea) {
  RHSEG_DISPATCH_K(K, {
    if (act_mode == RHSEG_ACT_SIGMOID)
      return fwd_upsampled<KK, RHSEG_ACT_SIGMOID>(z_lo, prev_probs, table, B, Hf, Wf, H, W, K_prev, logits, probs, psum, st, ea);
    if (act_mode == RHSEG_ACT_GROUPED)
      return fwd_upsampled<KK, RHSEG_ACT_GROUPED>(z_lo, prev_probs, table, B, Hf, Wf, H, W, K_prev, logits, probs, psum, st, ea);
    return fwd_upsampled<KK, RHSEG_ACT_ZEROS>(z_lo, prev_probs, table, B, Hf, Wf, H, W, K_prev, logits, probs, psum, st, ea);
  });
  return RHSEG_OK;
}

}  // namespace rhseg

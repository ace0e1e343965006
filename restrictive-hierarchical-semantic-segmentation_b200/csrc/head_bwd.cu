// Head backward: activation backward (composition / restrictive softmax / sigmoid), adjoint
// of the align_corners bilinear upsample, 1x1-conv backward (dfeats + fp64 weight-gradient
// sums) and the tiny parameter-gradient kernel (head conv, FiLM linear, uniform gradient to
// the previous level's pooled probabilities).  The reference obtains all of this from
// autograd over Models/models.py:58-77 and :263-306 / :757-802; closed forms in DESIGN.md.
#include <algorithm>
#include <cstdlib>
#include "common.cuh"
#include "pipeline.cuh"

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
  pdl_wait();
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

  float pp[K][VEC], dpar[K][VEC];
  if constexpr (MODE == RHSEG_ACT_GROUPED) {
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
  }
#pragma unroll
  for (int v = 0; v < VEC; ++v) {
    float zz[K], dPv[K], ppv[K], dzv[K], dparv[K];
#pragma unroll
    for (int k = 0; k < K; ++k) { zz[k] = z[k][v]; dPv[k] = dP[k][v]; ppv[k] = pp[k][v]; dzv[k] = dz[k][v]; }
    act_dz_pixel<K, MODE>(zz, dPv, ppv, li.start_mask, dzv, dparv);
#pragma unroll
    for (int k = 0; k < K; ++k) { dz[k][v] = dzv[k]; dpar[k][v] = dparv[k]; }
  }
  if constexpr (MODE == RHSEG_ACT_GROUPED) {
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
// 1x1 conv backward at feature resolution: TMA-fed, warp-specialised, persistent.
//
//   dfeats[b,c,n] = sum_k w[b,k,c] dz[b,k,n]            (pure write stream, registers -> global)
//   S[b,k,c]      = sum_n dz[b,k,n] feats[b,c,n]        (read stream + reduction)
//
// Work unit = (sample, channel stage of PIPE_CH channels, pixel tile), flattened in that order
// and split EVENLY over one resident wave of CTAs.  The producer warp streams the feature
// stage of each unit into the shared-memory ring (pipeline.cuh).  A consumer thread owns J
// vectors of VEC pixels: it re-loads dz for the tile (small, L2-resident) into registers,
// writes dfeats for the stage's channels, and forms K partial dot products per channel that
// are reduced across the warp with the transposed recursive-halving shuffle pattern
// (KP-1 + 5-log2(KP) shuffles).  The per-channel sums of the current (sample, stage) segment
// stay in registers across tiles and leave the warp as fp64 atomics when the segment ends.
// ------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------
// Parameter gradients as the TAIL of the conv backward (no second launch, no launch gap on the level chain):
// every CTA takes a ticket when its share of S / s has been added; the last one forms, from the complete sums,
//   d_head_w = sum_b (S_b diag(gamma_b) + s_b beta_b^T),  d_head_b = sum_b s_b,
//   dgamma_b[c] = sum_k S_b[k,c] W[k,c],  dbeta_b = W^T s_b,  d_film_w = sum_b [dgamma|dbeta]_b cond_b^T,
//   d_film_b = sum_b [dgamma|dbeta]_b,  g_prev[b] = film_w^T [dgamma|dbeta]_b
// (closed forms: DESIGN.md 3.5).  Same arguments as rhseg_head_param_grads; all sums in fp64.
// ------------------------------------------------------------------------------------
struct ParamTail {
  unsigned* ticket;  // device counter, zero on entry, left zero; nullptr = no tail
  const float* head_w;
  const float* film_w;
  const float* gamma_beta;
  const double* prev_psum;
  double n_pix;
  int B, K_prev;
  float* d_head_w;
  float* d_head_b;
  float* d_film_w;
  float* d_film_b;
  double* g_prev;
};

template <int K>
__device__ __noinline__ void param_grads_tail(const ParamTail& pt, const double* __restrict__ S, const double* __restrict__ s, int C,
                                              double* sm /* shared: B*K_prev + 2*B*C doubles */, int tid, int nthr) {
  const int B = pt.B, Kp = pt.film_w ? pt.K_prev : 0;
  const int lane = tid & 31, warp = tid >> 5, nwarp = nthr >> 5;
  double* cond = sm;                    // [B][Kp]  pooled probabilities of the previous level (fp32 values, as in the forward)
  double* dg = sm + (size_t)B * Kp;     // [B][C]   dgamma
  double* db = dg + (size_t)B * C;      // [B][C]   dbeta
  for (int i = tid; i < B * Kp; i += nthr) cond[i] = (double)(float)fast_div(pt.prev_psum[i], pt.n_pix);
  consumer_sync(nthr);
  // phase 1: thread <-> channel; every load of a channel is independent of the others (no barrier inside)
  for (int c = tid; c < C; c += nthr) {
    float w[K];
#pragma unroll
    for (int k = 0; k < K; ++k) w[k] = pt.head_w[(size_t)k * C + c];
    double dw[K], dfb_g = 0.0, dfb_b = 0.0, dfw_g[RHSEG_MAX_K], dfw_b[RHSEG_MAX_K];
#pragma unroll
    for (int k = 0; k < K; ++k) dw[k] = 0.0;
#pragma unroll
    for (int j = 0; j < RHSEG_MAX_K; ++j) { dfw_g[j] = 0.0; dfw_b[j] = 0.0; }
    for (int b = 0; b < B; ++b) {
      const double gam = Kp ? (double)pt.gamma_beta[(size_t)b * 2 * C + c] : 1.0;
      const double bet = Kp ? (double)pt.gamma_beta[(size_t)b * 2 * C + C + c] : 0.0;
      double dgam = 0.0, dbet = 0.0;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        const double Sv = __ldcg(S + ((size_t)b * K + k) * C + c);  // accumulated by other CTAs' atomics: read at L2
        const double sv = __ldcg(s + b * K + k);
        dw[k] += Sv * gam + sv * bet;
        dgam += Sv * (double)w[k];
        dbet += (double)w[k] * sv;
      }
      if (Kp) {
        dfb_g += dgam;
        dfb_b += dbet;
        dg[(size_t)b * C + c] = dgam;
        db[(size_t)b * C + c] = dbet;
#pragma unroll
        for (int j = 0; j < RHSEG_MAX_K; ++j)
          if (j < Kp) {
            dfw_g[j] += dgam * cond[b * Kp + j];
            dfw_b[j] += dbet * cond[b * Kp + j];
          }
      }
    }
#pragma unroll
    for (int k = 0; k < K; ++k) pt.d_head_w[(size_t)k * C + c] = (float)dw[k];
    if (Kp) {
      pt.d_film_b[c] = (float)dfb_g;
      pt.d_film_b[C + c] = (float)dfb_b;
#pragma unroll
      for (int j = 0; j < RHSEG_MAX_K; ++j)
        if (j < Kp) {
          pt.d_film_w[(size_t)c * Kp + j] = (float)dfw_g[j];
          pt.d_film_w[(size_t)(C + c) * Kp + j] = (float)dfw_b[j];
        }
    }
  }
  if (tid < K) {
    double acc = 0.0;
    for (int b = 0; b < B; ++b) acc += __ldcg(s + b * K + tid);
    pt.d_head_b[tid] = (float)acc;
  }
  if (Kp == 0) return;
  consumer_sync(nthr);
  // phase 2: g_prev[b][j] = sum_c film_w[c][j] dgamma[b][c] + film_w[C+c][j] dbeta[b][c]: one warp per (b, j) pair
  for (int p = warp; p < B * Kp; p += nwarp) {
    const int b = p / Kp, j = p - b * Kp;
    double acc = 0.0;
    for (int c = lane; c < C; c += 32)
      acc += (double)pt.film_w[(size_t)c * Kp + j] * dg[(size_t)b * C + c] + (double)pt.film_w[(size_t)(C + c) * Kp + j] * db[(size_t)b * C + c];
    acc = warp_sum(acc);
    if (lane == 0) pt.g_prev[p] = acc;
  }
}

// min-blocks 1 (0 = unspecified) for the scalar-row instances other than K = 2 / K = 4: left alone, ptxas squeezes them to
// 96 registers (+ spills) for a second CTA per SM that the shared-memory ring rules out anyway -- K = 3: 121 vs 112 us in
// the step.  K = 4 (140 registers unprompted) and K = 2 measured no better with the hint and keep ptxas' own choice.
template <int K, int VEC, int J, typename CFG>
__global__ void __launch_bounds__(CFG::THREADS, ((VEC == 1 && K != 2 && K != 4) ? 1 : 0))
conv_bwd_kernel(const float* __restrict__ feats, const float* __restrict__ dz, const float* __restrict__ eff_w,
                int C, int N, int n_tiles, int n_stages, int group, long units_total, int a0, int w_shared,
                float* __restrict__ dfeats, double* __restrict__ S, double* __restrict__ s, ParamTail pt) {
  pdl_wait();
  constexpr int KP = pad_k(K);
  // Register arrays and inner loops run over KC channels.  K = 3 computes on its zero-padded fourth channel (dz loads as
  // 0, the weight table is zero-padded already): ptxas schedules the K = 3 instance at 96 registers with little load
  // batching (111 us per HRNet launch against 96 us for K = 4 and K = 2); with four channels it IS the K = 4 schedule.
  constexpr int KC = (K == 3) ? 4 : K;
  constexpr int P = J * VEC;
  constexpr int NCONS = CFG::CONSUMERS;
  constexpr int T = NCONS * P;
  constexpr int ROWP = T + 4;
  constexpr int CH = CFG::CH;
  constexpr int NS = CFG::NS;
  constexpr int FLUSH_TILES = 64;  // fp32 running sums are handed to fp64 at least this often
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);  // full[NS], empty[NS]
  float* ring = reinterpret_cast<float*>(smem_raw + 128);  // [NS][CH][ROWP]
  float* w_all = ring + (size_t)NS * CH * ROWP;            // [NCW][CH][KP] per-warp weights of the current stage
  static_assert(2 * NS * 8 <= 128, "barrier block");
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long u_begin = units_total * blockIdx.x / gridDim.x;
  const long u_end = units_total * (blockIdx.x + 1) / gridDim.x;

  if (tid == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(smem_u32(&bars[i]), CFG::NPW);
      mbar_init(smem_u32(&bars[NS + i]), CFG::NCW);
    }
    mbar_fence_init();
  }
  __syncthreads();

  // Unit order: sample b -> GROUP of `group` consecutive pixel tiles -> channel stage -> tile of the group.
  // group >= n_tiles is the plain order (b, stage, tile).  Large planes x large batches (UNet 1024^2, batch 8..64) use
  // small groups: dz of a tile is read once per channel stage, and in the plain order those re-reads are a whole plane
  // apart -- 16 MB per sample, hundreds of MB over the CTAs in flight, i.e. from DRAM (+23 % traffic); inside a group
  // they are `group` tiles apart and hit L2.  Decoded once, then advanced incrementally.
  int stage, tile, b, g_first, g_cnt;
  {
    const long upb = (long)n_stages * n_tiles;  // units per sample
    b = (int)(u_begin / upb);
    const long r = u_begin - (long)b * upb;
    const long fgu = (long)n_stages * group;    // units of a full group
    const int gi = (int)(r / fgu);
    g_first = gi * group;
    g_cnt = min(group, n_tiles - g_first);
    const long r2 = r - (long)gi * fgu;
    stage = (int)(r2 / g_cnt);
    tile = g_first + (int)(r2 - (long)stage * g_cnt);
  }
  auto step_unit = [&](int& tile_, int& stage_, int& b_, int& gf_, int& gc_) {
    if (++tile_ == gf_ + gc_) {            // the group's tiles are done for this stage
      if (++stage_ == n_stages) {          // ... and for every stage: next group
        stage_ = 0;
        gf_ += gc_;
        if (gf_ == n_tiles) { gf_ = 0; ++b_; }
        gc_ = min(group, n_tiles - gf_);
      }
      tile_ = gf_;
    }
  };
  auto advance = [&]() { step_unit(tile, stage, b, g_first, g_cnt); };

  if (warp >= CFG::NCW) {
    // ------------------------------ producers ------------------------------
    if (lane == 0) {
      const int pw = warp - CFG::NCW;
      const uint64_t pol = l2_evict_first_policy();
      int slot = 0;
      uint32_t phase = 1;
      for (long u = u_begin; u < u_end; ++u) {
        mbar_wait(smem_u32(&bars[NS + slot]), phase);
        const int p0 = tile * T;
        const int c0 = stage * CH;
        issue_stage_rows(feats, ((long)b * C + c0) * N + p0, N, min(T, N - p0), min(CH, C - c0), pw, CFG::NPW, a0,
                         smem_u32(ring + (size_t)slot * CH * ROWP), ROWP * 4, smem_u32(&bars[slot]), pol);
        if (++slot == NS) { slot = 0; phase ^= 1u; }
        advance();
      }
    }
    return;
  }

  // -------------------------------- consumers --------------------------------
  const bool writer = transposed_writer<KP>(lane);
  const int my_k = transposed_index<KP>(lane);
  const int sN = N & 3;
  float* w_s = w_all + (size_t)warp * CH * KP;
  float sacc[CH];  // writer lanes: running sum over pixels of dz[my_k] * feats[c0 + cc]
  float s_acc = 0.f;
#pragma unroll
  for (int cc = 0; cc < CH; ++cc) sacc[cc] = 0.f;
  long cur_seg = -1;
  int seg_b = 0, seg_c0 = 0, seg_cnt = 0, tiles_since_flush = 0;

  auto flush = [&]() {  // per warp, no CTA-wide sync needed: every warp owns its partial sums
    if (writer && my_k < K) {
#pragma unroll
      for (int cc = 0; cc < CH; ++cc)
        if (cc < seg_cnt) atomicAdd(&S[((size_t)seg_b * K + my_k) * C + seg_c0 + cc], (double)sacc[cc]);
      if (seg_c0 == 0) atomicAdd(&s[seg_b * K + my_k], (double)s_acc);
    }
#pragma unroll
    for (int cc = 0; cc < CH; ++cc) sacc[cc] = 0.f;
    s_acc = 0.f;
    tiles_since_flush = 0;
  };

  float g_next[KC][P];
  auto load_dz = [&](int bb, int tt) {
    const int q0 = tt * T;
    const int q_act = min(T, N - q0);
#pragma unroll
    for (int j = 0; j < J; ++j) {
      const int l = (j * NCONS + tid) * VEC;
#pragma unroll
      for (int k = 0; k < KC; ++k) {
        Vec<VEC> t;
        if (k < K && l < q_act) t = ld_cached<VEC>(dz + ((size_t)bb * K + k) * N + q0 + l);
        else {
#pragma unroll
          for (int v = 0; v < VEC; ++v) t.v[v] = 0.f;
        }
#pragma unroll
        for (int v = 0; v < VEC; ++v) g_next[k][j * VEC + v] = t.v[v];
      }
    }
  };
  if (u_begin < u_end) load_dz(b, tile);
  int slot = 0;
  uint32_t phase = 0;
  for (long u = u_begin; u < u_end; ++u) {
    const long seg = ((long)b * n_tiles + g_first) * n_stages + stage;  // (sample, tile group, channel stage)
    const int p0 = tile * T;
    const int t_act = min(T, N - p0);
    const int c0 = stage * CH;
    const int ccnt = min(CH, C - c0);
    if (seg != cur_seg || tiles_since_flush >= FLUSH_TILES) {
      if (cur_seg >= 0) flush();
      if (seg != cur_seg) {  // effective weights of the stage's channels -> this warp's table [CH][KP]
        __syncwarp();
        for (int e = lane; e < CH * KP; e += 32) {
          const int cc = e / KP, k = e - cc * KP;
          w_s[e] = (cc < ccnt && k < K) ? __ldg(eff_w + ((size_t)(w_shared ? 0 : b) * K + k) * C + c0 + cc) : 0.f;
        }
        __syncwarp();
      }
      cur_seg = seg; seg_b = b; seg_c0 = c0; seg_cnt = ccnt;
    }
    ++tiles_since_flush;

    // dz of this tile was prefetched during the previous unit; fetch the NEXT unit's dz now so
    // that its L2 latency hides behind this unit's barrier wait and math
    int lp[J];
    bool ok[J];
    float g[KC][P];
#pragma unroll
    for (int j = 0; j < J; ++j) {
      lp[j] = (j * NCONS + tid) * VEC;
      ok[j] = lp[j] < t_act;
#pragma unroll
      for (int k = 0; k < KC; ++k)
#pragma unroll
        for (int v = 0; v < VEC; ++v) g[k][j * VEC + v] = g_next[k][j * VEC + v];
    }
    if (u + 1 < u_end) {
      int nt = tile, ns_ = stage, nb = b, ngf = g_first, ngc = g_cnt;
      step_unit(nt, ns_, nb, ngf, ngc);
      load_dz(nb, nt);
    }

    int off[4] = {0, 0, 0, 0};
    int sh0 = 0;
    if constexpr (VEC == 1) {
      sh0 = row_shift(((long)b * C + c0) * N + p0, a0);
#pragma unroll
      for (int q = 0; q < 4; ++q) off[q] = ((sh0 + q * sN) & 3) + tid;
    }
    float* dfb = dfeats ? dfeats + ((size_t)b * C + c0) * N + p0 : nullptr;
    mbar_wait(smem_u32(&bars[slot]), phase);
    const float* stage_base = ring + (size_t)slot * CH * ROWP;

    // `part` for pixels beyond the tile's end is harmless: their dz registers are zero
    auto channel = [&](int cc, int offv, float& acc_out) {
      const float* row = stage_base + cc * ROWP;
      float w[KP];
      if constexpr (KP == 4) {
        const float4 t = *reinterpret_cast<const float4*>(w_s + cc * KP);
        w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
      } else if constexpr (KP == 8) {
        const float4 t0 = *reinterpret_cast<const float4*>(w_s + cc * KP);
        const float4 t1 = *reinterpret_cast<const float4*>(w_s + cc * KP + 4);
        w[0] = t0.x; w[1] = t0.y; w[2] = t0.z; w[3] = t0.w; w[4] = t1.x; w[5] = t1.y; w[6] = t1.z; w[7] = t1.w;
      } else if constexpr (KP == 2) {
        const float2 t = *reinterpret_cast<const float2*>(w_s + cc * KP);
        w[0] = t.x; w[1] = t.y;
      } else {
        w[0] = w_s[cc];
      }
      float part[KP];
#pragma unroll
      for (int k = 0; k < KP; ++k) part[k] = 0.f;
#pragma unroll
      for (int j = 0; j < J; ++j) {
        float f[VEC];
        if constexpr (VEC == 4) {
          const float4 t = *reinterpret_cast<const float4*>(row + lp[j]);
          f[0] = t.x; f[1] = t.y; f[2] = t.z; f[3] = t.w;
        } else {
          f[0] = row[offv + j * NCONS];
        }
#pragma unroll
        for (int k = 0; k < KC; ++k)
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            // NaN-safe masking: stale shared memory past the tile end may hold any bit pattern
            const float fv = ok[j] ? f[v] : 0.f;
            part[k] = fmaf(g[k][j * VEC + v], fv, part[k]);
          }
        if (dfb && ok[j]) {
          Vec<VEC> o;
#pragma unroll
          for (int v = 0; v < VEC; ++v) {
            float a = 0.f;
#pragma unroll
            for (int k = 0; k < KC; ++k) a = fmaf(w[k], g[k][j * VEC + v], a);
            o.v[v] = a;
          }
          st_stream<VEC>(dfb + (size_t)cc * N + lp[j], o);
        }
      }
      warp_reduce_transposed<KP>(part, lane);
      acc_out += part[0];
    };
    if (ccnt == CH) {
#pragma unroll
      for (int cc = 0; cc < CH; ++cc) channel(cc, off[cc & 3], sacc[cc]);
    } else {
#pragma unroll
      for (int cc = 0; cc < CH; ++cc)
        if (cc < ccnt) channel(cc, ((sh0 + cc * sN) & 3) + tid, sacc[cc]);
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(smem_u32(&bars[NS + slot]));

    if (c0 == 0) {  // s[b][k] = sum_n dz: only the first channel stage contributes
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
    if (++slot == NS) { slot = 0; phase ^= 1u; }
    advance();
  }
  if (cur_seg >= 0) flush();
  // Level WITHOUT FiLM (level 0): its parameter gradients are plain sums over the batch, d_head_w = sum_b S_b and
  // d_head_b = sum_b s_b.  The last CTA to finish forms them here (a few thousand L2 loads) instead of a second kernel
  // at the very end of the step.  (The FiLM'd levels keep the separate 12-CTA kernel: their tail is the slower form.)
  if (pt.ticket != nullptr && pt.film_w == nullptr) {
    __shared__ int last_cta;
    __threadfence();
    consumer_sync(NCONS);
    if (tid == 0) last_cta = (atomicAdd(pt.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    consumer_sync(NCONS);
    if (last_cta) {
      __threadfence();
      for (int i = tid; i < K * C; i += NCONS) {
        double a = 0.0;
        for (int bb = 0; bb < pt.B; ++bb) a += __ldcg(S + (size_t)bb * K * C + i);  // accumulated by other CTAs: read at L2
        pt.d_head_w[i] = (float)a;
      }
      if (tid < K) {
        double a = 0.0;
        for (int bb = 0; bb < pt.B; ++bb) a += __ldcg(s + bb * K + tid);
        pt.d_head_b[tid] = (float)a;
      }
      if (tid == 0) *pt.ticket = 0u;  // ready for the next launch / graph replay
    }
  }
#ifdef RHSEG_WITH_PARAM_TAIL
  // Compiled out by default: measured slower than the separate 12-CTA kernel (25 us against 8.6 us at C = 720), and the
  // call into the tail raised the 128-bit instance of this kernel from 96 to 128 registers, i.e. from two CTAs per SM
  // to one (UNet conv backward 0.81 -> 0.75 of the HBM peak).
  if (pt.ticket != nullptr && pt.film_w != nullptr) {
    // last CTA done: S / s are complete -> parameter gradients (the ring is free: its memory holds the pool-gradient sums)
    __shared__ int is_last;
    __threadfence();
    consumer_sync(NCONS);
    if (tid == 0) is_last = (atomicAdd(pt.ticket, 1u) == gridDim.x - 1) ? 1 : 0;
    consumer_sync(NCONS);
    if (is_last) {
      __threadfence();
      param_grads_tail<K>(pt, S, s, C, reinterpret_cast<double*>(ring), tid, NCONS);
      if (tid == 0) *pt.ticket = 0u;  // ready for the next launch / graph replay
    }
  }
#endif
}

template <int K, int VEC, int J, typename CFG>
static int launch_conv_bwd(const float* feats, const float* dz, const float* eff_w, int B, int C, int N,
                           float* dfeats, double* S, double* s, int sm_count, cudaStream_t st, int w_shared,
                           const ParamTail& pt) {
  constexpr int KP = pad_k(K);
  constexpr int T = CFG::CONSUMERS * J * VEC;
  const size_t smem = 128 + ((size_t)CFG::NS * CFG::CH * (T + 4) + (size_t)CFG::NCW * CFG::CH * KP) * sizeof(float);
  auto kern = conv_bwd_kernel<K, VEC, J, CFG>;
  int per_sm = 0;
  RHSEG_CUDA(cached_launch_prep(reinterpret_cast<const void*>(kern), CFG::THREADS, smem, smem, &per_sm));
  if (per_sm < 1) return RHSEG_ERR_UNSUPPORTED;
  const int n_tiles = (N + T - 1) / T;
  const int n_stages = (C + CFG::CH - 1) / CFG::CH;
  const long units_total = (long)B * n_stages * n_tiles;
  // tile groups (see the kernel): only when dz no longer sits in L2 between two channel stages
  static int tune_group = -1;
  if (tune_group < 0) { const char* e = getenv("RHSEG_TUNE_BWD_GROUP"); tune_group = e ? atoi(e) : 0; }
  int group = n_tiles;
  if (tune_group > 0) group = std::min(n_tiles, tune_group);
  // measured, UNet tl 1024^2 batch 8 (dz = 134 MB), step in ms: plain 2.930 | groups of 2: 2.791 | 4: 2.656 | 6: 2.685 |
  // 8: 2.773 | 12: 2.908 | 16: 2.923; 620^2 batch 4 (dz = 25 MB, L2-resident anyway): plain 0.5507 | 4: 0.5557 | 8: 0.5519
  else if ((size_t)B * K * N * sizeof(float) > ((size_t)40 << 20) && n_stages > 1) group = std::min(n_tiles, 4);
  const long grid = std::max<long>(1, std::min<long>((long)sm_count * per_sm, units_total));
  const int a0 = (int)((reinterpret_cast<uintptr_t>(feats) >> 2) & 3);
  launch_pdl(kern, dim3((unsigned)grid), dim3(CFG::THREADS), smem, st, feats, dz, eff_w, C, N, n_tiles, n_stages, group, units_total, a0, w_shared, dfeats, S, s, pt);
  RHSEG_LAUNCH_CHECK();
  return RHSEG_OK;
}

static int tune_env(const char* name) {
  const char* v = getenv(name);
  return v ? atoi(v) : 0;
}
using BwdCfgV4 = PipeCfg<8, 8, 3>;  // 8 consumer warps x float4 -> T = 1024 px, 4 KB row copies, 33 KB stages
using BwdCfgS1 = PipeCfg<8, 16, 2>;  // scalar rows: T = 256*J px, 66 KB stages (dz reused over 16 channels)

// ------------------------------------------------------------------------------------
// parameter gradients (tiny): one thread per feature channel
// ------------------------------------------------------------------------------------
// Thread (cl, bl): channel blockIdx.x*PG_CH + cl, samples b = bl, bl+PG_BL, ...  All global loads of a
// thread are independent (one sample per thread when B <= PG_BL), the sums over samples and over
// channels go through shared memory.
constexpr int PG_CH = 64, PG_BL = 4, PG_THREADS = PG_CH * PG_BL;

__global__ void __launch_bounds__(PG_THREADS)
param_grads_kernel(const double* __restrict__ S, const double* __restrict__ s, const float* __restrict__ head_w,
                   const float* __restrict__ film_w, const float* __restrict__ gamma_beta,
                   const double* __restrict__ prev_psum, double n_pix, int B, int C, int K, int K_prev,
                   float* __restrict__ d_head_w, float* __restrict__ d_head_b, float* __restrict__ d_film_w,
                   float* __restrict__ d_film_b, double* __restrict__ g_prev) {
  pdl_wait();
  constexpr int NV = RHSEG_KERNEL_MAX_K + 2 + 2 * RHSEG_MAX_K;  // dw[k], dfb_g, dfb_b, dfw_g[j], dfw_b[j]
  __shared__ float red[PG_BL][PG_CH][NV + 1];
  __shared__ float gsm[PG_THREADS / 32][RHSEG_MAX_K];
  const int cl = threadIdx.x % PG_CH, bl = threadIdx.x / PG_CH;
  const int c = blockIdx.x * PG_CH + cl;
  const bool ok = c < C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (blockIdx.x == 0 && threadIdx.x < K) {
    double acc = 0.0;
    for (int b = 0; b < B; ++b) acc += s[b * K + threadIdx.x];
    d_head_b[threadIdx.x] = (float)acc;
  }
  float w[RHSEG_KERNEL_MAX_K], fwg[RHSEG_MAX_K], fwb[RHSEG_MAX_K];
#pragma unroll
  for (int k = 0; k < RHSEG_KERNEL_MAX_K; ++k) w[k] = (ok && k < K) ? head_w[(size_t)k * C + c] : 0.f;
#pragma unroll
  for (int j = 0; j < RHSEG_MAX_K; ++j) {
    fwg[j] = (ok && film_w && j < K_prev) ? film_w[(size_t)c * K_prev + j] : 0.f;
    fwb[j] = (ok && film_w && j < K_prev) ? film_w[(size_t)(C + c) * K_prev + j] : 0.f;
  }
  float acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.f;
  for (int b = bl; b < ((B + PG_BL - 1) / PG_BL) * PG_BL; b += PG_BL) {  // uniform trip count for the block reductions
    const bool live = ok && b < B;
    double dgam = 0.0, dbet = 0.0;
    if (live) {
      const double gam = film_w ? (double)gamma_beta[(size_t)b * 2 * C + c] : 1.0;
      const double bet = film_w ? (double)gamma_beta[(size_t)b * 2 * C + C + c] : 0.0;
#pragma unroll
      for (int k = 0; k < RHSEG_KERNEL_MAX_K; ++k)
        if (k < K) {
          const double Sv = S[((size_t)b * K + k) * C + c], sv = s[b * K + k];
          acc[k] += (float)(Sv * gam + sv * bet);
          dgam += Sv * (double)w[k];
          dbet += (double)w[k] * sv;
        }
    }
    if (film_w) {
      acc[RHSEG_KERNEL_MAX_K] += (float)dgam;
      acc[RHSEG_KERNEL_MAX_K + 1] += (float)dbet;
      for (int j = 0; j < K_prev; ++j) {
        const double cond = b < B ? (double)(float)fast_div(prev_psum[b * K_prev + j], n_pix) : 0.0;
        acc[RHSEG_KERNEL_MAX_K + 2 + j] += (float)(dgam * cond);
        acc[RHSEG_KERNEL_MAX_K + 2 + RHSEG_MAX_K + j] += (float)(dbet * cond);
        // g_prev[b][j] = sum over channels of film_w^T [dgamma|dbeta]: warp sums now, one block step below
        const float contrib = warp_sum((float)((double)fwg[j] * dgam + (double)fwb[j] * dbet));
        if (lane == 0) gsm[warp][j] = contrib;
      }
      __syncthreads();
      if (cl < K_prev && b < B) {  // thread (cl = j, bl): the PG_CH/32 warps that share this sample slot
        double tsum = 0.0;
        for (int q = 0; q < PG_CH / 32; ++q) tsum += (double)gsm[bl * (PG_CH / 32) + q][cl];
        atomicAdd(&g_prev[b * K_prev + cl], tsum);
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) red[bl][cl][i] = acc[i];
  __syncthreads();
  if (bl != 0 || !ok) return;
  float tot[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    tot[i] = 0.f;
#pragma unroll
    for (int q = 0; q < PG_BL; ++q) tot[i] += red[q][cl][i];
  }
#pragma unroll
  for (int k = 0; k < RHSEG_KERNEL_MAX_K; ++k)
    if (k < K) d_head_w[(size_t)k * C + c] = tot[k];
  if (film_w) {
    d_film_b[c] = tot[RHSEG_KERNEL_MAX_K];
    d_film_b[C + c] = tot[RHSEG_KERNEL_MAX_K + 1];
#pragma unroll
    for (int j = 0; j < RHSEG_MAX_K; ++j)
      if (j < K_prev) {
        d_film_w[(size_t)c * K_prev + j] = tot[RHSEG_KERNEL_MAX_K + 2 + j];
        d_film_w[(size_t)(C + c) * K_prev + j] = tot[RHSEG_KERNEL_MAX_K + 2 + RHSEG_MAX_K + j];
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
        launch_pdl(act_bwd_kernel<KK, 4, RHSEG_ACT_SIGMOID, THREADS>, dim3(grid), dim3(THREADS), 0, st, logits, prev_probs, table, dz_in, g_uniform, inv, dp_pix, pix_mask, K_prev, N, dz_out, dp_prev);
      else if (act_mode == RHSEG_ACT_GROUPED)
        launch_pdl(act_bwd_kernel<KK, 4, RHSEG_ACT_GROUPED, THREADS>, dim3(grid), dim3(THREADS), 0, st, logits, prev_probs, table, dz_in, g_uniform, inv, dp_pix, pix_mask, K_prev, N, dz_out, dp_prev);
      else
        launch_pdl(act_bwd_kernel<KK, 4, RHSEG_ACT_ZEROS, THREADS>, dim3(grid), dim3(THREADS), 0, st, logits, prev_probs, table, dz_in, g_uniform, inv, dp_pix, pix_mask, K_prev, N, dz_out, dp_prev);
    } else {
      dim3 grid((unsigned)((N + THREADS - 1) / THREADS), B);
      if (act_mode == RHSEG_ACT_SIGMOID)
        launch_pdl(act_bwd_kernel<KK, 1, RHSEG_ACT_SIGMOID, THREADS>, dim3(grid), dim3(THREADS), 0, st, logits, prev_probs, table, dz_in, g_uniform, inv, dp_pix, pix_mask, K_prev, N, dz_out, dp_prev);
      else if (act_mode == RHSEG_ACT_GROUPED)
        launch_pdl(act_bwd_kernel<KK, 1, RHSEG_ACT_GROUPED, THREADS>, dim3(grid), dim3(THREADS), 0, st, logits, prev_probs, table, dz_in, g_uniform, inv, dp_pix, pix_mask, K_prev, N, dz_out, dp_prev);
      else
        launch_pdl(act_bwd_kernel<KK, 1, RHSEG_ACT_ZEROS, THREADS>, dim3(grid), dim3(THREADS), 0, st, logits, prev_probs, table, dz_in, g_uniform, inv, dp_pix, pix_mask, K_prev, N, dz_out, dp_prev);
    }
  });
  RHSEG_LAUNCH_CHECK();
  return RHSEG_OK;
}

static int conv_bwd_impl(const float* feats, const float* dz, const float* eff_w, int B, int C, int K, int n_pix,
                         float* dfeats, double* S, double* s, int zero_sums, void* stream, const ParamTail& pt) {
  if (!feats || !dz || !eff_w || !S || !s || B <= 0 || C <= 0 || n_pix <= 0) return RHSEG_ERR_ARG;
  if (K < 1 || K > RHSEG_KERNEL_MAX_K) return RHSEG_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  const int w_shared = (zero_sums & 2) ? 1 : 0;
  if (zero_sums & 1) {
    RHSEG_CUDA(cudaMemsetAsync(S, 0, sizeof(double) * (size_t)B * K * C, st));
    RHSEG_CUDA(cudaMemsetAsync(s, 0, sizeof(double) * (size_t)B * K, st));
  }
  const int sms = device_sm_count();
  RHSEG_DISPATCH_K(K, {
    const bool v4 = (n_pix % 4 == 0) && ((reinterpret_cast<uintptr_t>(feats) & 15u) == 0) &&
                    ((reinterpret_cast<uintptr_t>(dz) & 15u) == 0) && ((reinterpret_cast<uintptr_t>(dfeats) & 15u) == 0);
    if (v4) {
      if constexpr (KK == 4) {
        const int t = tune_env("RHSEG_TUNE_BWD_V4");
        if (t == 1) return launch_conv_bwd<KK, 4, 1, PipeCfg<4, 8, 3>>(feats, dz, eff_w, B, C, n_pix, dfeats, S, s, sms, st, w_shared, pt);
        if (t == 2) return launch_conv_bwd<KK, 4, 1, PipeCfg<8, 16, 2>>(feats, dz, eff_w, B, C, n_pix, dfeats, S, s, sms, st, w_shared, pt);
        if (t == 3) return launch_conv_bwd<KK, 4, 1, PipeCfg<8, 16, 3>>(feats, dz, eff_w, B, C, n_pix, dfeats, S, s, sms, st, w_shared, pt);
      }
      return launch_conv_bwd<KK, 4, 1, BwdCfgV4>(feats, dz, eff_w, B, C, n_pix, dfeats, S, s, sms, st, w_shared, pt);
    }
    if constexpr (KK == 4) {
      const int t = tune_env("RHSEG_TUNE_BWD_S1");
      if (t == 1) return launch_conv_bwd<KK, 1, 4, PipeCfg<8, 16, 3>>(feats, dz, eff_w, B, C, n_pix, dfeats, S, s, sms, st, w_shared, pt);
      if (t == 2) return launch_conv_bwd<KK, 1, 4, PipeCfg<8, 24, 2>>(feats, dz, eff_w, B, C, n_pix, dfeats, S, s, sms, st, w_shared, pt);
      if (t == 3) return launch_conv_bwd<KK, 1, 4, PipeCfg<8, 16, 2>>(feats, dz, eff_w, B, C, n_pix, dfeats, S, s, sms, st, w_shared, pt);
    }
    return launch_conv_bwd<KK, 1, (KK <= 4 ? 4 : 2), BwdCfgS1>(feats, dz, eff_w, B, C, n_pix, dfeats, S, s, sms, st, w_shared, pt);
  });
  return RHSEG_OK;
}

extern "C" int rhseg_head_conv_bwd(const float* feats, const float* dz, const float* eff_w, int B, int C, int K,
                                   int n_pix, float* dfeats, double* S, double* s, int zero_sums, void* stream) {
  ParamTail none{};
  return conv_bwd_impl(feats, dz, eff_w, B, C, K, n_pix, dfeats, S, s, zero_sums, stream, none);
}

extern "C" int rhseg_head_conv_bwd_params(const float* feats, const float* dz, const float* eff_w, int B, int C, int K,
                                          int n_pix, float* dfeats, double* S, double* s, int flags, const float* head_w,
                                          const float* film_w, const float* gamma_beta, const double* prev_psum,
                                          double n_pix_out, int K_prev, float* d_head_w, float* d_head_b, float* d_film_w,
                                          float* d_film_b, double* g_prev, unsigned* ticket, void* stream) {
  if (!head_w || !d_head_w || !d_head_b || !ticket) return RHSEG_ERR_ARG;
  if (film_w) {
    if (!gamma_beta || !prev_psum || !d_film_w || !d_film_b || !g_prev || n_pix_out <= 0) return RHSEG_ERR_ARG;
    if (K_prev < 1 || K_prev > RHSEG_MAX_K) return RHSEG_ERR_UNSUPPORTED;
  }
  // Measured (HRNet-W48 tl, B = 4): the single-CTA tail costs 25 us per launch against 8.6 us for the 12-CTA
  // rhseg_head_param_grads kernel behind a programmatic dependent launch, so the two-kernel form is the default;
  // RHSEG_PARAM_TAIL=1 selects the tail (kept for small C / B where one CTA is enough).
  static int use_tail = -1;
#ifdef RHSEG_WITH_PARAM_TAIL
  if (use_tail < 0) { const char* e = getenv("RHSEG_PARAM_TAIL"); use_tail = (e && e[0] == '1') ? 1 : 0; }
#else
  use_tail = 0;
#endif
  // level 0 of a narrow donor: the batch sums are formed by the conv kernel's last CTA.  Measured: UNet (K*C*B = 1024 sums)
  // 0.5507 vs 0.5538 ms per step with the separate kernel; HRNet-W48 (11520 sums through one CTA) 0.4207 vs 0.4166 -> kept
  // for small heads only.
  if (!film_w && (long)K * C * B <= 4096 && !getenv("RHSEG_NO_L0_TAIL")) {
    ParamTail pt{ticket, head_w, nullptr, nullptr, nullptr, n_pix_out, B, 0, d_head_w, d_head_b, nullptr, nullptr, nullptr};
    return conv_bwd_impl(feats, dz, eff_w, B, C, K, n_pix, dfeats, S, s, flags, stream, pt);
  }
  if (!use_tail || (film_w && ((long)B * K_prev + 2L * B * C) * 8 > 60 * 1024)) {
    ParamTail none{};
    const int rc = conv_bwd_impl(feats, dz, eff_w, B, C, K, n_pix, dfeats, S, s, flags, stream, none);
    if (rc != RHSEG_OK) return rc;
    // g_prev: the caller's buffer is zero on entry when bit 2 (value 4) of flags is set (part of a larger zero fill)
    return rhseg_head_param_grads(S, s, head_w, film_w, gamma_beta, prev_psum, n_pix_out, B, C, K, K_prev, d_head_w, d_head_b,
                                  d_film_w, d_film_b, g_prev, (flags & 4) ? 1 : 0, stream);
  }
  ParamTail pt{ticket, head_w, film_w, gamma_beta, prev_psum, n_pix_out, B, K_prev, d_head_w, d_head_b, d_film_w, d_film_b, g_prev};
  return conv_bwd_impl(feats, dz, eff_w, B, C, K, n_pix, dfeats, S, s, flags, stream, pt);
}

extern "C" int rhseg_head_param_grads(const double* S, const double* s, const float* head_w, const float* film_w,
                                      const float* gamma_beta, const double* prev_psum, double n_pix, int B, int C,
                                      int K, int K_prev, float* d_head_w, float* d_head_b, float* d_film_w,
                                      float* d_film_b, double* g_prev, int g_prev_zeroed, void* stream) {
  if (!S || !s || !head_w || !d_head_w || !d_head_b || B <= 0 || C <= 0) return RHSEG_ERR_ARG;
  if (K < 1 || K > RHSEG_KERNEL_MAX_K) return RHSEG_ERR_UNSUPPORTED;
  cudaStream_t st = (cudaStream_t)stream;
  if (film_w) {
    if (!gamma_beta || !prev_psum || !d_film_w || !d_film_b || !g_prev || n_pix <= 0) return RHSEG_ERR_ARG;
    if (K_prev < 1 || K_prev > RHSEG_MAX_K) return RHSEG_ERR_UNSUPPORTED;
    if (!g_prev_zeroed) RHSEG_CUDA(cudaMemsetAsync(g_prev, 0, sizeof(double) * (size_t)B * K_prev, st));
  }
  launch_pdl(param_grads_kernel, dim3((C + PG_CH - 1) / PG_CH), dim3(PG_THREADS), 0, st, S, s, head_w, film_w, gamma_beta, prev_psum, n_pix, B, C, K, K_prev,
                                                      d_head_w, d_head_b, d_film_w, d_film_b, g_prev);
  RHSEG_LAUNCH_CHECK();
  return RHSEG_OK;
}

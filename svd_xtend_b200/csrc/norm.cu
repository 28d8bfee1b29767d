// GroupNorm(32)(+SiLU) and LayerNorm, forward and backward, over channels-last bf16 activations.
// HBM-bound CUDA-core kernels: coalesced bf16x2 row reads (one warp per row, lanes stride the
// channel axis), fp32 statistics, warp-shuffle / shared-memory reductions, fp32 atomics for the
// cross-CTA partial sums.
//
// Replaces F.group_norm + F.silu of ResnetBlock2D / TemporalResnetBlock / TransformerSpatioTemporalModel
// [D: diffusers models/resnet.py, transformer_temporal.py] and conv_norm_out
// (/root/reference/src/unet_spatio_temporal_condition.py:238-239,480-481), and F.layer_norm of
// BasicTransformerBlock / TemporalBasicTransformerBlock [D: models/attention.py].
#include "common.cuh"
#include "../../include/svd_xtend_b200.h"
#include "host_util.h"

namespace svdx {

constexpr int GN_THREADS = 256;
constexpr int GN_WARPS = GN_THREADS / 32;
constexpr int GN_ROWS_PER_CTA = 64;
constexpr int MAX_C = 2560;  // channel pairs are kept in smem accumulators: 2 * MAX_C floats

struct GnSrc {
  const bf16* x;
  long long ldx;
  int C1;
  const bf16* x2;
  long long ldx2;
  int C2;
};

SVDX_DEVINL float2 load_pair(const GnSrc& s, long long row, int c) {
  const bf16* p = (c < s.C1) ? (s.x + row * s.ldx + c) : (s.x2 + row * s.ldx2 + (c - s.C1));
  return unpack_bf16x2(*reinterpret_cast<const uint32_t*>(p));
}

// ------------------------------------------------------------------ GroupNorm statistics
// grid (row_chunks, outer). Accumulates sum / sumsq into mean[] / rstd[] (pre-zeroed), finalised below.
__global__ void __launch_bounds__(GN_THREADS) gn_stats_partial(GnSrc s, int rows, int G, float* sum, float* sumsq) {
  __shared__ float sh_s[32 * 2];
  const int C = s.C1 + s.C2;
  const int cpg = C / G;
  const int n = blockIdx.y;
  const int r0 = blockIdx.x * GN_ROWS_PER_CTA;
  const int r1 = min(r0 + GN_ROWS_PER_CTA, rows);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 2 * G; i += GN_THREADS) sh_s[i] = 0.f;
  __syncthreads();
  // a lane owns channel pairs c = 2*lane + 64*j; accumulate per pair over the warp's rows, flush per j
  // blockIdx.z selects a 64-channel slab; a lane owns the channel pair c = slab*64 + 2*lane over the warp's rows
  const int c = blockIdx.z * 64 + 2 * lane;
  if (c < C) {
    float a = 0.f, b = 0.f;
#pragma unroll 4
    for (int r = r0 + warp; r < r1; r += GN_WARPS) {
      const float2 v = load_pair(s, (long long)n * rows + r, c);
      a += v.x + v.y;
      b += v.x * v.x + v.y * v.y;
    }
    const int g = c / cpg;
    atomicAdd(&sh_s[g], a);
    atomicAdd(&sh_s[G + g], b);
  }
  __syncthreads();
  const int g_lo = (blockIdx.z * 64) / cpg, g_hi = min(C - 1, blockIdx.z * 64 + 63) / cpg;
  for (int i = g_lo + threadIdx.x; i <= g_hi; i += GN_THREADS) {
    atomicAdd(&sum[n * G + i], sh_s[i]);
    atomicAdd(&sumsq[n * G + i], sh_s[G + i]);
  }
}

__global__ void gn_stats_finalize(float* mean, float* rstd, int total, float inv_count, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float m = mean[i] * inv_count;
  const float var = fmaxf(rstd[i] * inv_count - m * m, 0.f);
  mean[i] = m;
  rstd[i] = rsqrtf(var + eps);
}

// ------------------------------------------------------------------ GroupNorm apply (+SiLU)
// one thread per 8 output channels (16 B store)
__global__ void __launch_bounds__(256) gn_apply_kernel(GnSrc s, long long total_rows, int rows, int G, const float* __restrict__ mean,
                                                       const float* __restrict__ rstd, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, int fuse_silu, bf16* __restrict__ y, long long ldy) {
  const int C = s.C1 + s.C2;
  const int vec_per_row = C / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total_rows * vec_per_row) return;
  const long long row = idx / vec_per_row;
  const int c0 = (int)(idx - row * vec_per_row) * 8;
  const int n = (int)(row / rows);
  const int cpg = C / G;
  const bf16* p = (c0 < s.C1) ? (s.x + row * s.ldx + c0) : (s.x2 + row * s.ldx2 + (c0 - s.C1));
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t in[4] = {u.x, u.y, u.z, u.w};
  uint32_t out[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + 2 * k;
    const int g = c / cpg;
    const float m = mean[n * G + g], rs = rstd[n * G + g];
    float2 v = unpack_bf16x2(in[k]);
    float a = (v.x - m) * rs * gamma[c] + beta[c];
    float b = (v.y - m) * rs * gamma[c + 1] + beta[c + 1];
    if (fuse_silu) { a = silu_f(a); b = silu_f(b); }
    out[k] = pack_bf16x2(a, b);
  }
  *reinterpret_cast<uint4*>(y + row * ldy + c0) = make_uint4(out[0], out[1], out[2], out[3]);
}

// ------------------------------------------------------------------ GroupNorm backward
// pass 1: per (n, group) s1 = sum(g*gamma), s2 = sum(g*gamma*xhat), g = dy * silu'(z); optional dgamma/dbeta
__global__ void __launch_bounds__(GN_THREADS) gn_bwd_partial(GnSrc s, const bf16* __restrict__ dy, long long lddy, int rows, int G,
                                                             const float* __restrict__ mean, const float* __restrict__ rstd,
                                                             const float* __restrict__ gamma, const float* __restrict__ beta,
                                                             int fuse_silu, float* ws, float* dgamma, float* dbeta) {
  __shared__ float sh_s[32 * 2];
  const int C = s.C1 + s.C2;
  const int cpg = C / G;
  const int n = blockIdx.y;
  const int r0 = blockIdx.x * GN_ROWS_PER_CTA;
  const int r1 = min(r0 + GN_ROWS_PER_CTA, rows);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 2 * G; i += GN_THREADS) sh_s[i] = 0.f;
  __syncthreads();
  const int c = blockIdx.z * 64 + 2 * lane;
  if (c < C) {
    const int g = c / cpg;
    const float m = mean[n * G + g], rs = rstd[n * G + g];
    const float g0 = gamma[c], g1 = gamma[c + 1], b0 = beta[c], b1 = beta[c + 1];
    float a1 = 0.f, a2 = 0.f;          // group sums
    float dg0 = 0.f, dg1 = 0.f, db0 = 0.f, db1 = 0.f;
#pragma unroll 4
    for (int r = r0 + warp; r < r1; r += GN_WARPS) {
      const long long row = (long long)n * rows + r;
      const float2 v = load_pair(s, row, c);
      const float2 d = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dy + row * lddy + c));
      const float xh0 = (v.x - m) * rs, xh1 = (v.y - m) * rs;
      float e0 = d.x, e1 = d.y;
      if (fuse_silu) {
        e0 *= silu_grad_f(xh0 * g0 + b0);
        e1 *= silu_grad_f(xh1 * g1 + b1);
      }
      a1 += e0 * g0 + e1 * g1;
      a2 += e0 * g0 * xh0 + e1 * g1 * xh1;
      dg0 += e0 * xh0; dg1 += e1 * xh1; db0 += e0; db1 += e1;
    }
    atomicAdd(&sh_s[g], a1);
    atomicAdd(&sh_s[G + g], a2);
    if (dgamma) {
      atomicAdd(&dgamma[c], dg0); atomicAdd(&dgamma[c + 1], dg1);
      atomicAdd(&dbeta[c], db0); atomicAdd(&dbeta[c + 1], db1);
    }
  }
  __syncthreads();
  const int g_lo = (blockIdx.z * 64) / cpg, g_hi = min(C - 1, blockIdx.z * 64 + 63) / cpg;
  for (int i = g_lo + threadIdx.x; i <= g_hi; i += GN_THREADS) {
    atomicAdd(&ws[(n * G + i) * 2 + 0], sh_s[i]);
    atomicAdd(&ws[(n * G + i) * 2 + 1], sh_s[G + i]);
  }
}

// pass 2: dx = rstd * (g*gamma - s1/cnt - xhat * s2/cnt)
__global__ void __launch_bounds__(256) gn_bwd_apply(GnSrc s, const bf16* __restrict__ dy, long long lddy, long long total_rows, int rows,
                                                    int G, const float* __restrict__ mean, const float* __restrict__ rstd,
                                                    const float* __restrict__ gamma, const float* __restrict__ beta, int fuse_silu,
                                                    const float* __restrict__ ws, float inv_count, bf16* __restrict__ dx, long long lddx,
                                                    bf16* __restrict__ dx2, long long lddx2) {
  const int C = s.C1 + s.C2;
  const int vec_per_row = C / 8;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total_rows * vec_per_row) return;
  const long long row = idx / vec_per_row;
  const int c0 = (int)(idx - row * vec_per_row) * 8;
  const int n = (int)(row / rows);
  const int cpg = C / G;
  const bool first = c0 < s.C1;
  const bf16* p = first ? (s.x + row * s.ldx + c0) : (s.x2 + row * s.ldx2 + (c0 - s.C1));
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint4 ud = *reinterpret_cast<const uint4*>(dy + row * lddy + c0);
  const uint32_t in[4] = {u.x, u.y, u.z, u.w};
  const uint32_t din[4] = {ud.x, ud.y, ud.z, ud.w};
  uint32_t out[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = c0 + 2 * k;
    const int g = c / cpg;
    const float m = mean[n * G + g], rs = rstd[n * G + g];
    const float s1 = ws[(n * G + g) * 2] * inv_count, s2 = ws[(n * G + g) * 2 + 1] * inv_count;
    const float2 v = unpack_bf16x2(in[k]);
    const float2 d = unpack_bf16x2(din[k]);
    const float xh0 = (v.x - m) * rs, xh1 = (v.y - m) * rs;
    float e0 = d.x, e1 = d.y;
    if (fuse_silu) {
      e0 *= silu_grad_f(xh0 * gamma[c] + beta[c]);
      e1 *= silu_grad_f(xh1 * gamma[c + 1] + beta[c + 1]);
    }
    const float o0 = rs * (e0 * gamma[c] - s1 - xh0 * s2);
    const float o1 = rs * (e1 * gamma[c + 1] - s1 - xh1 * s2);
    out[k] = pack_bf16x2(o0, o1);
  }
  bf16* q = first ? (dx + row * lddx + c0) : (dx2 + row * lddx2 + (c0 - s.C1));
  *reinterpret_cast<uint4*>(q) = make_uint4(out[0], out[1], out[2], out[3]);
}

// ------------------------------------------------------------------ LayerNorm
constexpr int LN_MAXJ = 40;  // C <= 2560

template <int NJ>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const bf16* __restrict__ x, long long ldx, int rows, int C,
                                                     const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                     bf16* __restrict__ y, long long ldy, float* __restrict__ mean, float* __restrict__ rstd,
                                                     const float* __restrict__ addvec, int add_div, bf16* __restrict__ xsum, long long ldxs) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const bf16* xr = x + (long long)warp * ldx;
  const float* av = addvec ? addvec + (long long)(warp / add_div) * C : nullptr;
  float2 v[NJ];
  float sum = 0.f;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int c = 2 * lane + 64 * j;
    if (c < C) {
      v[j] = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(xr + c));
      if (av) {
        // x + addvec is rounded to bf16 (it is the residual stream value the reference materialises)
        const uint32_t pk = pack_bf16x2(v[j].x + av[c], v[j].y + av[c + 1]);
        *reinterpret_cast<uint32_t*>(xsum + (long long)warp * ldxs + c) = pk;
        v[j] = unpack_bf16x2(pk);
      }
      sum += v[j].x + v[j].y;
    } else {
      v[j] = make_float2(0.f, 0.f);
    }
  }
  sum = warp_sum(sum);
  const float m = sum / C;
  float sq = 0.f;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int c = 2 * lane + 64 * j;
    if (c < C) {
      const float a = v[j].x - m, b = v[j].y - m;
      sq += a * a + b * b;
    }
  }
  sq = warp_sum(sq);
  const float rs = rsqrtf(sq / C + eps);
  if (lane == 0) { mean[warp] = m; rstd[warp] = rs; }
  bf16* yr = y + (long long)warp * ldy;
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int c = 2 * lane + 64 * j;
    if (c < C) {
      const float a = (v[j].x - m) * rs * gamma[c] + beta[c];
      const float b = (v[j].y - m) * rs * gamma[c + 1] + beta[c + 1];
      *reinterpret_cast<uint32_t*>(yr + c) = pack_bf16x2(a, b);
    }
  }
}

// backward: each warp walks rows (grid-stride), keeps per-lane dgamma/dbeta partials in registers,
// flushes them with atomics at the end. dx = rstd*(gamma*dy - mean(gamma*dy) - xhat*mean(gamma*dy*xhat)) [+ dres]
template <int NJ>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const bf16* __restrict__ x, long long ldx, const bf16* __restrict__ dy, long long lddy,
                                                     int rows, int C, const float* __restrict__ gamma, const float* __restrict__ mean,
                                                     const float* __restrict__ rstd, bf16* __restrict__ dx, long long lddx,
                                                     const bf16* __restrict__ dres, long long lddres, float* dgamma, float* dbeta) {
  const int lane = threadIdx.x & 31;
  const int warp0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  float2 gm[NJ], dg[NJ], db[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const int c = 2 * lane + 64 * j;
    gm[j] = (c < C) ? make_float2(gamma[c], gamma[c + 1]) : make_float2(0.f, 0.f);
    dg[j] = make_float2(0.f, 0.f);
    db[j] = make_float2(0.f, 0.f);
  }
  for (int row = warp0; row < rows; row += nwarps) {
    const bf16* xr = x + (long long)row * ldx;
    const bf16* dr = dy + (long long)row * lddy;
    const float m = mean[row], rs = rstd[row];
    float2 xh[NJ], gd[NJ];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = 2 * lane + 64 * j;
      if (c < C) {
        const float2 v = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(xr + c));
        const float2 d = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dr + c));
        xh[j] = make_float2((v.x - m) * rs, (v.y - m) * rs);
        gd[j] = make_float2(d.x * gm[j].x, d.y * gm[j].y);
        s1 += gd[j].x + gd[j].y;
        s2 += gd[j].x * xh[j].x + gd[j].y * xh[j].y;
        dg[j].x += d.x * xh[j].x; dg[j].y += d.y * xh[j].y;
        db[j].x += d.x; db[j].y += d.y;
      } else {
        xh[j] = make_float2(0.f, 0.f);
        gd[j] = make_float2(0.f, 0.f);
      }
    }
    s1 = warp_sum(s1) / C;
    s2 = warp_sum(s2) / C;
    bf16* oxr = dx + (long long)row * lddx;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = 2 * lane + 64 * j;
      if (c < C) {
        float a = rs * (gd[j].x - s1 - xh[j].x * s2);
        float b = rs * (gd[j].y - s1 - xh[j].y * s2);
        if (dres) {
          const float2 r = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(dres + (long long)row * lddres + c));
          a += r.x; b += r.y;
        }
        *reinterpret_cast<uint32_t*>(oxr + c) = pack_bf16x2(a, b);
      }
    }
  }
  if (dgamma) {
    // block-level reduction through shared memory, then one atomic per channel per CTA
    __shared__ float sh_g[64 * NJ], sh_b[64 * NJ];
    for (int i = threadIdx.x; i < 64 * NJ; i += blockDim.x) { sh_g[i] = 0.f; sh_b[i] = 0.f; }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = 2 * lane + 64 * j;
      if (c < C) {
        atomicAdd(&sh_g[c], dg[j].x); atomicAdd(&sh_g[c + 1], dg[j].y);
        atomicAdd(&sh_b[c], db[j].x); atomicAdd(&sh_b[c + 1], db[j].y);
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
      atomicAdd(&dgamma[i], sh_g[i]);
      atomicAdd(&dbeta[i], sh_b[i]);
    }
  }
}

}  // namespace svdx

using namespace svdx;

static int gn_check(int C1, int C2, int G, int64_t ldx, int64_t ldx2, const void* x, const void* x2) {
  const int C = C1 + C2;
  if (!x || G <= 0 || G > 32 || C % G || (C / G) % 2 || C % 8 || C1 % 8 || C2 % 8 || C > 2 * MAX_C) return 1;
  if (C2 > 0 && !x2) return 1;
  if (ldx % 8 || (C2 > 0 && ldx2 % 8)) return 1;
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (x2 && (reinterpret_cast<uintptr_t>(x2) & 15))) return 1;
  return 0;
}

extern "C" int svdx_groupnorm_stats(const void* x, int64_t ldx, int32_t C1, const void* x2, int64_t ldx2, int32_t C2, int32_t outer,
                                    int32_t rows, int32_t num_groups, float eps, float* mean, float* rstd, void* stream_v) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_v);
  if (gn_check(C1, C2, num_groups, ldx, ldx2, x, x2) || outer <= 0 || rows <= 0 || !mean || !rstd)
    return svdx_fail(SVDX_E_BADARG, "groupnorm_stats: bad arguments");
  GnSrc s{reinterpret_cast<const bf16*>(x), ldx, C1, reinterpret_cast<const bf16*>(x2), ldx2, C2};
  const int total = outer * num_groups;
  cudaMemsetAsync(mean, 0, sizeof(float) * total, st);
  cudaMemsetAsync(rstd, 0, sizeof(float) * total, st);
  dim3 grid((rows + GN_ROWS_PER_CTA - 1) / GN_ROWS_PER_CTA, outer, (C1 + C2 + 63) / 64);
  gn_stats_partial<<<grid, GN_THREADS, 0, st>>>(s, rows, num_groups, mean, rstd);
  const float inv = 1.0f / ((float)rows * (float)((C1 + C2) / num_groups));
  gn_stats_finalize<<<(total + 127) / 128, 128, 0, st>>>(mean, rstd, total, inv, eps);
  SVDX_CHECK_LAUNCH("groupnorm_stats");
  return SVDX_OK;
}

extern "C" int svdx_groupnorm_apply(const void* x, int64_t ldx, int32_t C1, const void* x2, int64_t ldx2, int32_t C2, int32_t outer,
                                    int32_t rows, int32_t num_groups, const float* mean, const float* rstd, const float* gamma,
                                    const float* beta, int32_t fuse_silu, void* y, int64_t ldy, void* stream_v) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_v);
  if (gn_check(C1, C2, num_groups, ldx, ldx2, x, x2) || !y || ldy % 8 || (reinterpret_cast<uintptr_t>(y) & 15) || !mean || !rstd || !gamma || !beta)
    return svdx_fail(SVDX_E_BADARG, "groupnorm_apply: bad arguments");
  GnSrc s{reinterpret_cast<const bf16*>(x), ldx, C1, reinterpret_cast<const bf16*>(x2), ldx2, C2};
  const long long total_rows = (long long)outer * rows;
  const long long nvec = total_rows * ((C1 + C2) / 8);
  gn_apply_kernel<<<(unsigned)((nvec + 255) / 256), 256, 0, st>>>(s, total_rows, rows, num_groups, mean, rstd, gamma, beta, fuse_silu,
                                                                 reinterpret_cast<bf16*>(y), ldy);
  SVDX_CHECK_LAUNCH("groupnorm_apply");
  return SVDX_OK;
}

extern "C" int svdx_groupnorm_bwd(const void* x, int64_t ldx, int32_t C1, const void* x2, int64_t ldx2, int32_t C2, const void* dy,
                                  int64_t lddy, int32_t outer, int32_t rows, int32_t num_groups, const float* mean, const float* rstd,
                                  const float* gamma, const float* beta, int32_t fuse_silu, void* dx, int64_t lddx, void* dx2,
                                  int64_t lddx2, float* dgamma, float* dbeta, float* workspace, void* stream_v) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_v);
  if (gn_check(C1, C2, num_groups, ldx, ldx2, x, x2) || !dy || lddy % 8 || !dx || lddx % 8 || (C2 > 0 && (!dx2 || lddx2 % 8)) || !workspace ||
      (dgamma && !dbeta))
    return svdx_fail(SVDX_E_BADARG, "groupnorm_bwd: bad arguments");
  GnSrc s{reinterpret_cast<const bf16*>(x), ldx, C1, reinterpret_cast<const bf16*>(x2), ldx2, C2};
  const int total = outer * num_groups;
  cudaMemsetAsync(workspace, 0, sizeof(float) * 2 * total, st);
  dim3 grid((rows + GN_ROWS_PER_CTA - 1) / GN_ROWS_PER_CTA, outer, (C1 + C2 + 63) / 64);
  gn_bwd_partial<<<grid, GN_THREADS, 0, st>>>(s, reinterpret_cast<const bf16*>(dy), lddy, rows, num_groups, mean, rstd, gamma, beta,
                                              fuse_silu, workspace, dgamma, dbeta);
  const long long total_rows = (long long)outer * rows;
  const long long nvec = total_rows * ((C1 + C2) / 8);
  const float inv = 1.0f / ((float)rows * (float)((C1 + C2) / num_groups));
  gn_bwd_apply<<<(unsigned)((nvec + 255) / 256), 256, 0, st>>>(s, reinterpret_cast<const bf16*>(dy), lddy, total_rows, rows, num_groups,
                                                              mean, rstd, gamma, beta, fuse_silu, workspace, inv,
                                                              reinterpret_cast<bf16*>(dx), lddx, reinterpret_cast<bf16*>(dx2), lddx2);
  SVDX_CHECK_LAUNCH("groupnorm_bwd");
  return SVDX_OK;
}

template <int NJ>
static void ln_fwd_launch(const void* x, int64_t ldx, int rows, int C, const float* g, const float* b, float eps, void* y, int64_t ldy,
                          float* mean, float* rstd, const float* addvec, int add_div, void* xsum, int64_t ldxs, cudaStream_t st) {
  const int warps_per_cta = 8;
  ln_fwd_kernel<NJ><<<(rows + warps_per_cta - 1) / warps_per_cta, 256, 0, st>>>(reinterpret_cast<const bf16*>(x), ldx, rows, C, g, b, eps,
                                                                              reinterpret_cast<bf16*>(y), ldy, mean, rstd, addvec, add_div,
                                                                              reinterpret_cast<bf16*>(xsum), ldxs);
}

extern "C" int svdx_layernorm_fwd(const void* x, int64_t ldx, int32_t rows, int32_t C, const float* gamma, const float* beta, float eps,
                                  void* y, int64_t ldy, float* mean, float* rstd, const float* addvec, int32_t add_div, void* xsum,
                                  int64_t ldxs, void* stream_v) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_v);
  if (!x || !y || !gamma || !beta || !mean || !rstd || rows <= 0 || C <= 0 || C % 2 || C > 64 * LN_MAXJ || ldx % 2 || ldy % 2 ||
      (addvec && (!xsum || add_div <= 0 || ldxs % 2)))
    return svdx_fail(SVDX_E_BADARG, "layernorm_fwd: bad arguments");
  const int nj = (C + 63) / 64;
  if (nj <= 5) ln_fwd_launch<5>(x, ldx, rows, C, gamma, beta, eps, y, ldy, mean, rstd, addvec, add_div, xsum, ldxs, st);
  else if (nj <= 10) ln_fwd_launch<10>(x, ldx, rows, C, gamma, beta, eps, y, ldy, mean, rstd, addvec, add_div, xsum, ldxs, st);
  else if (nj <= 20) ln_fwd_launch<20>(x, ldx, rows, C, gamma, beta, eps, y, ldy, mean, rstd, addvec, add_div, xsum, ldxs, st);
  else ln_fwd_launch<40>(x, ldx, rows, C, gamma, beta, eps, y, ldy, mean, rstd, addvec, add_div, xsum, ldxs, st);
  SVDX_CHECK_LAUNCH("layernorm_fwd");
  return SVDX_OK;
}

template <int NJ>
static void ln_bwd_launch(const void* x, int64_t ldx, const void* dy, int64_t lddy, int rows, int C, const float* g, const float* mean,
                          const float* rstd, void* dx, int64_t lddx, const void* dres, int64_t lddres, float* dgamma, float* dbeta,
                          cudaStream_t st) {
  int ctas = (rows + 7) / 8;
  const int cap = svdx_num_sms() * 2;   // few CTAs: the dgamma/dbeta atomics contend per address
  if (ctas > cap) ctas = cap;
  ln_bwd_kernel<NJ><<<ctas, 256, 0, st>>>(reinterpret_cast<const bf16*>(x), ldx, reinterpret_cast<const bf16*>(dy), lddy, rows, C, g, mean,
                                          rstd, reinterpret_cast<bf16*>(dx), lddx, reinterpret_cast<const bf16*>(dres), lddres, dgamma, dbeta);
}

extern "C" int svdx_layernorm_bwd(const void* x, int64_t ldx, const void* dy, int64_t lddy, int32_t rows, int32_t C, const float* gamma,
                                  const float* mean, const float* rstd, void* dx, int64_t lddx, const void* dres, int64_t lddres,
                                  float* dgamma, float* dbeta, void* stream_v) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream_v);
  if (!x || !dy || !dx || !gamma || !mean || !rstd || rows <= 0 || C <= 0 || C % 2 || C > 64 * LN_MAXJ || ldx % 2 || lddy % 2 || lddx % 2 ||
      (dgamma && !dbeta) || (dres && lddres % 2))
    return svdx_fail(SVDX_E_BADARG, "layernorm_bwd: bad arguments");
  const int nj = (C + 63) / 64;
  if (nj <= 5) ln_bwd_launch<5>(x, ldx, dy, lddy, rows, C, gamma, mean, rstd, dx, lddx, dres, lddres, dgamma, dbeta, st);
  else if (nj <= 10) ln_bwd_launch<10>(x, ldx, dy, lddy, rows, C, gamma, mean, rstd, dx, lddx, dres, lddres, dgamma, dbeta, st);
  else if (nj <= 20) ln_bwd_launch<20>(x, ldx, dy, lddy, rows, C, gamma, mean, rstd, dx, lddx, dres, lddres, dgamma, dbeta, st);
  else ln_bwd_launch<40>(x, ldx, dy, lddy, rows, C, gamma, mean, rstd, dx, lddx, dres, lddres, dgamma, dbeta, st);
  SVDX_CHECK_LAUNCH("layernorm_bwd");
  return SVDX_OK;
}

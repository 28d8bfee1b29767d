// Short-sequence attention (S <= 32 tokens, head_dim 64): the temporal self-attention of the SVD UNet
// (TemporalBasicTransformerBlock.attn1 [D], one sequence of T = 14 / 25 frames per pixel and head), forward and backward.
//
// Why a second kernel family: this op is HBM-bound, not tensor-bound — per (pixel, head) it does 2 x 14 x 14 x 64 MACs on
// 3 x 14 x 128 bytes of input (0.64 GFLOP over 92 MB at the 40x64 level: 7 FLOP/byte). The tcgen05 kernel of attention.cu
// has to pack 8 pixel sequences into one 128-row tile and mask 7/8 of the 128 x 128 scores; every CTA pays barrier set-up,
// a TMEM allocation and three TMA round trips for 6 KB of useful data and ran at 12 % (forward) / 8 x off (backward) of
// the HBM roofline (profiles/r2_kbench.txt: 56 us / 234 us at the 40x64 level against 14 us / 28 us of pure traffic).
// Here one warp owns SPW sequences of one head (two when S <= 16): K and V (backward: also Q and dO) are staged once in
// shared memory as fp32 rows, lane t is query t (forward, backward phase A) or key t (backward phase B), all arithmetic is
// fp32 FMA, exp2 on the MUFU, no atomics, grid-stride over (sequence group, head). Token addressing follows the SvdxAttn
// descriptor (outer / inner / token strides), so the strided "frames of one pixel" gather needs no permute.
//
// forward : S = Q K^T * scale -> softmax over the S keys -> O = P V,  lse = log sum exp (natural log, as attention.cu)
// backward: P recomputed; dP = dO V^T; delta = sum_s P dP; dS = P (dP - delta) * scale;
//           dQ = dS K (phase A, lane = query);  dV = P^T dO, dK = dS^T Q (phase B, lane = key, P / dS through shared memory)
#include "common.cuh"
#include "../../include/svd_xtend_b200.h"
#include "host_util.h"
#include <stdlib.h>

namespace svdx {

constexpr float AS_LOG2E = 1.4426950408889634f;
constexpr int AS_PAD = 33;   // row stride of the per-warp [32][S] score scratch (conflict-free column reads)

struct SmallAttnP {
  const bf16 *q, *k, *v, *o, *dout;
  bf16 *out, *dq, *dk, *dv;
  long long ldq, ldk, ldv, ldo, lddo, lddq, lddk, lddv;
  long long nseq;
  int heads, S, inner;
  long long outer_stride, inner_stride, tok_stride;
  float scale;
  float* lse;
};

SVDX_DEVINL long long as_token(const SmallAttnP& p, long long seq, int t) {
  const long long outer = seq / p.inner;
  const long long i = seq - outer * p.inner;
  return outer * p.outer_stride + i * p.inner_stride + (long long)t * p.tok_stride;
}

// 8 bf16 (one 16-byte vector) -> 8 floats in shared memory
SVDX_DEVINL void as_store8(float* dst, const uint4 u) {
  const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  *reinterpret_cast<float4*>(dst) = make_float4(a.x, a.y, b.x, b.y);
  *reinterpret_cast<float4*>(dst + 4) = make_float4(c.x, c.y, d.x, d.y);
}

// stage the [S][64] head slices of SPW sequences: row r = sq * S + t, 8 lanes per 128-byte row
template <int SPW>
SVDX_DEVINL void as_stage(const SmallAttnP& p, const bf16* __restrict__ src, long long ld, long long sg, int h, int lane, float* dst) {
  const int S = p.S;
  for (int i = lane; i < SPW * S * 8; i += 32) {
    const int r = i >> 3, c = i & 7;
    const int sq = r / S, t = r - sq * S;
    const long long seq = sg * SPW + sq;
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    if (seq < p.nseq) u = __ldg(reinterpret_cast<const uint4*>(src + as_token(p, seq, t) * ld + h * 64 + c * 8));
    as_store8(dst + r * 64 + c * 8, u);
  }
}

// one token's 64-wide head slice into registers (scaled)
SVDX_DEVINL void as_load_row(const bf16* __restrict__ src, float (&f)[64], float s) {
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(src) + c);
    const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), cc = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
    f[8 * c] = a.x * s; f[8 * c + 1] = a.y * s; f[8 * c + 2] = b.x * s; f[8 * c + 3] = b.y * s;
    f[8 * c + 4] = cc.x * s; f[8 * c + 5] = cc.y * s; f[8 * c + 6] = d.x * s; f[8 * c + 7] = d.y * s;
  }
}
SVDX_DEVINL void as_store_row(bf16* dst, const float (&f)[64], float s) {
#pragma unroll
  for (int c = 0; c < 8; ++c)
    reinterpret_cast<uint4*>(dst)[c] = make_uint4(pack_bf16x2(f[8 * c] * s, f[8 * c + 1] * s), pack_bf16x2(f[8 * c + 2] * s, f[8 * c + 3] * s),
                                                  pack_bf16x2(f[8 * c + 4] * s, f[8 * c + 5] * s), pack_bf16x2(f[8 * c + 6] * s, f[8 * c + 7] * s));
}
// dot of a register row with a shared-memory row (broadcast reads), 4 independent chains
SVDX_DEVINL float as_dot(const float (&f)[64], const float* __restrict__ row) {
  float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float4 r = *reinterpret_cast<const float4*>(row + 4 * j);
    d0 = fmaf(f[4 * j], r.x, d0); d1 = fmaf(f[4 * j + 1], r.y, d1); d2 = fmaf(f[4 * j + 2], r.z, d2); d3 = fmaf(f[4 * j + 3], r.w, d3);
  }
  return (d0 + d1) + (d2 + d3);
}
SVDX_DEVINL void as_axpy(float (&acc)[64], float a, const float* __restrict__ row) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float4 r = *reinterpret_cast<const float4*>(row + 4 * j);
    acc[4 * j] = fmaf(a, r.x, acc[4 * j]); acc[4 * j + 1] = fmaf(a, r.y, acc[4 * j + 1]);
    acc[4 * j + 2] = fmaf(a, r.z, acc[4 * j + 2]); acc[4 * j + 3] = fmaf(a, r.w, acc[4 * j + 3]);
  }
}

constexpr int ASF_WARPS = 4;
template <int SPW>
__global__ void __launch_bounds__(ASF_WARPS * 32) attn_small_fwd_kernel(const SmallAttnP p) {
  extern __shared__ float as_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.S;
  float* sK = as_smem + warp * (2 * SPW * S * 64 + 32 * AS_PAD);
  float* sV = sK + SPW * S * 64;
  float* sc = sV + SPW * S * 64 + lane * AS_PAD;
  const long long ngroups = (p.nseq + SPW - 1) / SPW;
  const long long total = ngroups * p.heads;
  const float qs = p.scale * AS_LOG2E;
  const int sq = (SPW == 2) ? (lane >> 4) : 0;
  const int t = (SPW == 2) ? (lane & 15) : lane;
  for (long long pi = (long long)blockIdx.x * ASF_WARPS + warp; pi < total; pi += (long long)gridDim.x * ASF_WARPS) {
    const long long sg = pi / p.heads;
    const int h = (int)(pi - sg * p.heads);
    as_stage<SPW>(p, p.k, p.ldk, sg, h, lane, sK);
    as_stage<SPW>(p, p.v, p.ldv, sg, h, lane, sV);
    __syncwarp();
    const long long seq = sg * SPW + sq;
    const bool act = (t < S) && (seq < p.nseq);
    const long long tok = act ? as_token(p, seq, t) : 0;
    const float* Kq = sK + sq * S * 64;
    const float* Vq = sV + sq * S * 64;
    float m = -INFINITY;
    {
      float qf[64];
      if (act) as_load_row(p.q + tok * p.ldq + h * 64, qf, qs);
      else {
#pragma unroll
        for (int i = 0; i < 64; ++i) qf[i] = 0.f;
      }
      for (int s = 0; s < S; ++s) {
        const float d = as_dot(qf, Kq + s * 64);
        sc[s] = d;
        m = fmaxf(m, d);
      }
    }
    float l = 0.f;
    for (int s = 0; s < S; ++s) {
      const float e = exp2_fast(sc[s] - m);
      sc[s] = e;
      l += e;
    }
    float o[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) o[i] = 0.f;
    for (int s = 0; s < S; ++s) as_axpy(o, sc[s], Vq + s * 64);
    if (act) {
      as_store_row(p.out + tok * p.ldo + h * 64, o, 1.0f / l);
      if (p.lse) p.lse[tok * p.heads + h] = m * 0.6931471805599453f + __logf(l);
    }
    __syncwarp();   // everyone is done with sK / sV before the next group overwrites them
  }
}

constexpr int ASB_WARPS = 2;
template <int SPW>
__global__ void __launch_bounds__(ASB_WARPS * 32) attn_small_bwd_kernel(const SmallAttnP p) {
  extern __shared__ float as_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.S;
  const int tile = SPW * S * 64;
  float* sK = as_smem + warp * (4 * tile + 2 * SPW * 32 * AS_PAD);
  float* sV = sK + tile;
  float* sQ = sV + tile;
  float* sD = sQ + tile;
  float* sP = sD + tile;                 // [SPW][32][AS_PAD]  probabilities P[t][s]
  float* sS = sP + SPW * 32 * AS_PAD;    // [SPW][32][AS_PAD]  dS[t][s]
  const long long ngroups = (p.nseq + SPW - 1) / SPW;
  const long long total = ngroups * p.heads;
  const float qs = p.scale * AS_LOG2E;
  const int sq = (SPW == 2) ? (lane >> 4) : 0;
  const int t = (SPW == 2) ? (lane & 15) : lane;
  for (long long pi = (long long)blockIdx.x * ASB_WARPS + warp; pi < total; pi += (long long)gridDim.x * ASB_WARPS) {
    const long long sg = pi / p.heads;
    const int h = (int)(pi - sg * p.heads);
    as_stage<SPW>(p, p.k, p.ldk, sg, h, lane, sK);
    as_stage<SPW>(p, p.v, p.ldv, sg, h, lane, sV);
    as_stage<SPW>(p, p.q, p.ldq, sg, h, lane, sQ);
    as_stage<SPW>(p, p.dout, p.lddo, sg, h, lane, sD);
    __syncwarp();
    const long long seq = sg * SPW + sq;
    const bool act = (t < S) && (seq < p.nseq);
    const long long tok = act ? as_token(p, seq, t) : 0;
    const float* Kq = sK + sq * tile / SPW;
    const float* Vq = sV + sq * tile / SPW;
    const float* Qq = sQ + sq * tile / SPW;
    const float* Dq = sD + sq * tile / SPW;
    float* pr = sP + (sq * 32 + t) * AS_PAD;     // row t of this sequence's P
    float* dr = sS + (sq * 32 + t) * AS_PAD;     // row t of dS
    // ---------------- phase A: lane = query t
    float m = -INFINITY;
    {
      float qf[64];
      if (act) as_load_row(p.q + tok * p.ldq + h * 64, qf, qs);
      else {
#pragma unroll
        for (int i = 0; i < 64; ++i) qf[i] = 0.f;
      }
      for (int s = 0; s < S; ++s) {
        const float d = as_dot(qf, Kq + s * 64);
        pr[s] = d;
        m = fmaxf(m, d);
      }
    }
    {
      float df[64];
      if (act) as_load_row(p.dout + tok * p.lddo + h * 64, df, 1.0f);
      else {
#pragma unroll
        for (int i = 0; i < 64; ++i) df[i] = 0.f;
      }
      for (int s = 0; s < S; ++s) dr[s] = as_dot(df, Vq + s * 64);     // dP[t][s]
    }
    float l = 0.f;
    for (int s = 0; s < S; ++s) {
      const float e = exp2_fast(pr[s] - m);
      pr[s] = e;
      l += e;
    }
    const float inv_l = 1.0f / l;
    float delta = 0.f;
    for (int s = 0; s < S; ++s) {
      const float pv = pr[s] * inv_l;
      pr[s] = pv;
      delta = fmaf(pv, dr[s], delta);
    }
    for (int s = 0; s < S; ++s) dr[s] = pr[s] * (dr[s] - delta) * p.scale;    // dS[t][s]
    {
      float dq[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) dq[i] = 0.f;
      for (int s = 0; s < S; ++s) as_axpy(dq, dr[s], Kq + s * 64);
      if (act) as_store_row(p.dq + tok * p.lddq + h * 64, dq, 1.0f);
    }
    __syncwarp();
    // ---------------- phase B: lane = key s (same (sq, t) mapping, t now indexes the key)
    const float* pcol = sP + sq * 32 * AS_PAD + t;    // P[tq][t]  = pcol[tq * AS_PAD]
    const float* dcol = sS + sq * 32 * AS_PAD + t;
    {
      float dv[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) dv[i] = 0.f;
      for (int tq = 0; tq < S; ++tq) as_axpy(dv, pcol[tq * AS_PAD], Dq + tq * 64);
      if (act) as_store_row(p.dv + tok * p.lddv + h * 64, dv, 1.0f);
    }
    {
      float dk[64];
#pragma unroll
      for (int i = 0; i < 64; ++i) dk[i] = 0.f;
      for (int tq = 0; tq < S; ++tq) as_axpy(dk, dcol[tq * AS_PAD], Qq + tq * 64);
      if (act) as_store_row(p.dk + tok * p.lddk + h * 64, dk, 1.0f);
    }
    __syncwarp();
  }
}

}  // namespace svdx

using namespace svdx;

// eligible: strided short sequences (the temporal attention); env SVDX_ATTN_SMALL=0 routes them back to the tcgen05 kernels
bool svdx_attention_small_eligible(const SvdxAttn* d) {
  static int on = -1;
  if (on < 0) { const char* e = getenv("SVDX_ATTN_SMALL"); on = (e && e[0] == '0') ? 0 : 1; }
  return on && d && d->inner > 1 && d->S > 0 && d->S <= 32;
}

static int small_fill(const SvdxAttn* d, SmallAttnP& p, bool bwd) {
  if (!d->q || !d->k || !d->v || !d->o) return svdx_fail(SVDX_E_BADARG, "attention(small): null pointer");
  if (d->heads <= 0 || d->nseq <= 0 || d->inner <= 0 || d->nseq % d->inner) return svdx_fail(SVDX_E_BADARG, "attention(small): bad sequence geometry");
  if ((d->ldq % 8) || (d->ldk % 8) || (d->ldv % 8) || (d->ldo % 8)) return svdx_fail(SVDX_E_BADARG, "attention(small): leading dims must be multiples of 8");
  const uintptr_t al = reinterpret_cast<uintptr_t>(d->q) | reinterpret_cast<uintptr_t>(d->k) | reinterpret_cast<uintptr_t>(d->v) | reinterpret_cast<uintptr_t>(d->o);
  if (al & 15) return svdx_fail(SVDX_E_BADARG, "attention(small): operands must be 16 B aligned");
  memset(&p, 0, sizeof(p));
  p.q = reinterpret_cast<const bf16*>(d->q); p.k = reinterpret_cast<const bf16*>(d->k); p.v = reinterpret_cast<const bf16*>(d->v);
  p.o = reinterpret_cast<const bf16*>(d->o); p.out = reinterpret_cast<bf16*>(d->o);
  p.ldq = d->ldq; p.ldk = d->ldk; p.ldv = d->ldv; p.ldo = d->ldo;
  p.nseq = d->nseq; p.heads = d->heads; p.S = d->S; p.inner = d->inner;
  p.outer_stride = d->outer_stride; p.inner_stride = d->inner_stride; p.tok_stride = d->tok_stride;
  p.scale = d->scale; p.lse = d->lse;
  if (bwd) {
    if (!d->dout || !d->dq || !d->dk || !d->dv) return svdx_fail(SVDX_E_BADARG, "attention_bwd(small): null pointer");
    if ((d->lddo % 8) || (d->lddq % 8) || (d->lddk % 8) || (d->lddv % 8)) return svdx_fail(SVDX_E_BADARG, "attention_bwd(small): leading dims");
    const uintptr_t a2 = reinterpret_cast<uintptr_t>(d->dout) | reinterpret_cast<uintptr_t>(d->dq) | reinterpret_cast<uintptr_t>(d->dk) | reinterpret_cast<uintptr_t>(d->dv);
    if (a2 & 15) return svdx_fail(SVDX_E_BADARG, "attention_bwd(small): operands must be 16 B aligned");
    p.dout = reinterpret_cast<const bf16*>(d->dout); p.lddo = d->lddo;
    p.dq = reinterpret_cast<bf16*>(d->dq); p.dk = reinterpret_cast<bf16*>(d->dk); p.dv = reinterpret_cast<bf16*>(d->dv);
    p.lddq = d->lddq; p.lddk = d->lddk; p.lddv = d->lddv;
  }
  return SVDX_OK;
}

template <typename K>
static int small_launch(K kernel, const SmallAttnP& p, int spw, int warps, size_t smem, cudaStream_t st, bool* attr_flags, const char* what) {
  const int slot = svdx_device_slot();
  if (!attr_flags[slot]) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return svdx_fail_cuda(e, what);
    attr_flags[slot] = true;
  }
  const long long total = ((p.nseq + spw - 1) / spw) * p.heads;
  long long ctas = (total + warps - 1) / warps;
  const long long cap = 8LL * svdx_num_sms();
  if (ctas > cap) ctas = cap;
  kernel<<<(unsigned)ctas, warps * 32, smem, st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return svdx_fail_cuda(e, what);
  return SVDX_OK;
}

int svdx_attention_small_fwd(const SvdxAttn* d, cudaStream_t st) {
  SmallAttnP p;
  int rc = small_fill(d, p, false);
  if (rc) return rc;
  const int spw = d->S <= 16 ? 2 : 1;
  const size_t smem = (size_t)ASF_WARPS * (2 * spw * d->S * 64 + 32 * AS_PAD) * sizeof(float);
  // the dynamic shared-memory limit depends on S: keep one flag per (device, spw) and always request the S = 32 / 16 maximum
  const size_t smem_max = (size_t)ASF_WARPS * (2 * spw * (spw == 2 ? 16 : 32) * 64 + 32 * AS_PAD) * sizeof(float);
  static bool f1[SVDX_MAX_DEVICES] = {false}, f2[SVDX_MAX_DEVICES] = {false};
  if (spw == 2) {
    const int slot = svdx_device_slot();
    if (!f2[slot]) {
      cudaError_t e = cudaFuncSetAttribute(attn_small_fwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
      if (e != cudaSuccess) return svdx_fail_cuda(e, "attention_fwd(small): smem attribute");
      f2[slot] = true;
    }
    return small_launch(attn_small_fwd_kernel<2>, p, 2, ASF_WARPS, smem, st, f2, "attention_fwd(small)");
  }
  const int slot = svdx_device_slot();
  if (!f1[slot]) {
    cudaError_t e = cudaFuncSetAttribute(attn_small_fwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
    if (e != cudaSuccess) return svdx_fail_cuda(e, "attention_fwd(small): smem attribute");
    f1[slot] = true;
  }
  return small_launch(attn_small_fwd_kernel<1>, p, 1, ASF_WARPS, smem, st, f1, "attention_fwd(small)");
}

int svdx_attention_small_bwd(const SvdxAttn* d, cudaStream_t st) {
  SmallAttnP p;
  int rc = small_fill(d, p, true);
  if (rc) return rc;
  const int spw = d->S <= 16 ? 2 : 1;
  const size_t smem = (size_t)ASB_WARPS * (4 * spw * d->S * 64 + 2 * spw * 32 * AS_PAD) * sizeof(float);
  const size_t smem_max = (size_t)ASB_WARPS * (4 * spw * (spw == 2 ? 16 : 32) * 64 + 2 * spw * 32 * AS_PAD) * sizeof(float);
  static bool f1[SVDX_MAX_DEVICES] = {false}, f2[SVDX_MAX_DEVICES] = {false};
  const int slot = svdx_device_slot();
  if (spw == 2) {
    if (!f2[slot]) {
      cudaError_t e = cudaFuncSetAttribute(attn_small_bwd_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
      if (e != cudaSuccess) return svdx_fail_cuda(e, "attention_bwd(small): smem attribute");
      f2[slot] = true;
    }
    return small_launch(attn_small_bwd_kernel<2>, p, 2, ASB_WARPS, smem, st, f2, "attention_bwd(small)");
  }
  if (!f1[slot]) {
    cudaError_t e = cudaFuncSetAttribute(attn_small_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max);
    if (e != cudaSuccess) return svdx_fail_cuda(e, "attention_bwd(small): smem attribute");
    f1[slot] = true;
  }
  return small_launch(attn_small_bwd_kernel<1>, p, 1, ASB_WARPS, smem, st, f1, "attention_bwd(small)");
}

// svdx_tapgemm, CTA-pair variant: tcgen05.mma.cta_group::2 on a 2-CTA cluster (two SMs of one TPC).
//
// Why: with one CTA per 128 x N tile every k-block pulls 16 KB of A + N*128 B of B through L2 for
// 128*N*64 MACs; at N = 160..256 that is ~0.1 KB/clk/SM, i.e. > 15 KB/clk chip-wide, about twice what
// L2 delivers — the single-CTA kernel tops out near 650 TFLOP/s (profiles/). A CTA pair computes a
// 256 x N tile: each CTA loads its own 128 rows of A and only HALF of the B tile (N/2 rows); the MMA unit
// reads the other half from the peer's shared memory. Same FLOPs, ~35-45 % less L2->SM traffic per FLOP.
//
// Protocol (mirrors the 1-CTA kernel, see tapgemm.cu, plus the cluster plumbing):
//   * both CTAs run a TMA producer; every load uses .cta_group::2 and signals the LEADER's (rank 0) full
//     barrier (peer bit of the mbarrier address cleared); the leader arms it with the bytes of both CTAs;
//   * only the leader issues tcgen05.mma.cta_group::2 (M = 256); tcgen05.commit ... multicast::cluster frees
//     the smem stage in BOTH CTAs and publishes the accumulator to BOTH epilogues;
//   * each CTA's epilogue drains its own 128 TMEM lanes and arrives (remotely for rank 1) on the leader's
//     tmem_empty barrier; cluster barriers fence set-up and teardown.
// Restricted to K-major operands without split-K (the weight-gradient forms stay on the 1-CTA kernel).
#include "tapgemm_common.cuh"
#include <stdlib.h>

namespace svdx {

constexpr int STAGES2 = 6;
constexpr int B2_STAGE_BYTES = 128 * BLOCK_K * 2;  // half of a <=256-row B tile
constexpr int SMEM2_BYTES = 1024 + STAGES2 * (A_STAGE_BYTES + B2_STAGE_BYTES) + NUM_EPI_WARPS * EPI_STAGE_BYTES + 256;
// WIDE variant: a 256 x 320 tile (block_n == 320) computed as two N = 160 MMAs per k-step on the same A stage. The A tile is
// pulled through L2 once per 320 output columns instead of once per 160 (C = 320 / 640 layers: -31 % L2 -> SM bytes per
// FLOP, the bound of these kernels). 320 fp32 columns cannot be double-buffered in 512 TMEM columns, so the two accumulators
// OVERLAP: even tiles use columns [0, 320), odd tiles [192, 512). The epilogue drains the shared columns [192, 320) of its
// tile first and then releases the next tile's MMAs, which run while the remaining columns are drained.
constexpr int STAGES2W = 5;
constexpr int B2W_STAGE_BYTES = 160 * BLOCK_K * 2;  // this CTA's half of a 320-row B tile: rows [0,80) -> MMA 0, [80,160) -> MMA 1
constexpr int SMEM2W_BYTES = 1024 + STAGES2W * (A_STAGE_BYTES + B2W_STAGE_BYTES) + NUM_EPI_WARPS * EPI_STAGE_BYTES + 256;
constexpr int WIDE_ACC1 = 192;                      // TMEM column base of the odd tiles' accumulator
constexpr uint32_t PEER_BIT_MASK = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address

SVDX_DEVINL uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
SVDX_DEVINL void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
SVDX_DEVINL void tma2_load_2d(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1)
      : "memory");
}
SVDX_DEVINL void tma2_load_3d(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
SVDX_DEVINL void tma2_load_4d(const CUtensorMap* m, uint32_t bar, uint32_t dst, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar & PEER_BIT_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
SVDX_DEVINL void tmem_alloc2(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
}
SVDX_DEVINL void tmem_relinquish2() { asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory"); }
SVDX_DEVINL void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
SVDX_DEVINL void umma2_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at the same smem offset in every CTA of `mask` once the prior MMAs retire
SVDX_DEVINL void umma2_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
// relaxed: the only hazard this arrive orders is the TMEM read-before-overwrite, which tcgen05.wait::ld + the tcgen05 fence
// cover; a .release.cluster arrive compiles to MEMBAR.ALL.GPU and stalls every tile on all outstanding global traffic.
SVDX_DEVINL void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}\n"
      ::"r"(bar), "r"(cta)
      : "memory");
}

template <int EPI, bool WIDE = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1) tapgemm2_kernel(const __grid_constant__ TapGemmKParams p) {
  constexpr int STAGES2 = WIDE ? svdx::STAGES2W : svdx::STAGES2;                 // shadow the namespace constants
  constexpr int B2_STAGE_BYTES = WIDE ? svdx::B2W_STAGE_BYTES : svdx::B2_STAGE_BYTES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sA = smem_base;
  const uint32_t sB = smem_base + STAGES2 * A_STAGE_BYTES;
  const uint32_t sEpi = sB + STAGES2 * B2_STAGE_BYTES;
  const uint32_t sBar = sEpi + NUM_EPI_WARPS * EPI_STAGE_BYTES;
  const uint32_t bar_full = sBar;
  const uint32_t bar_empty = sBar + 8 * STAGES2;
  const uint32_t bar_tfull = sBar + 16 * STAGES2;
  const uint32_t bar_tempty = bar_tfull + 8 * ACC_STAGES;
  const uint32_t tmem_slot = bar_tempty + 8 * ACC_STAGES;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&p.tma);
    prefetch_tmap(&p.tmb);
    for (int i = 0; i < STAGES2; ++i) {
      mbar_init(bar_full + 8 * i, 1);
      mbar_init(bar_empty + 8 * i, 1);
    }
    for (int i = 0; i < ACC_STAGES; ++i) {
      mbar_init(bar_tfull + 8 * i, 1);
      mbar_init(bar_tempty + 8 * i, 2 * NUM_EPI_WARPS);  // epilogue warps of BOTH CTAs (used on the leader only)
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc2(tmem_slot, TMEM_COLS);
    tmem_relinquish2();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;

  const int total_tiles = p.pair_m_tiles * p.n_tiles;
  const int bn_out = p.geglu ? p.block_n / 2 : p.block_n;
  const int b_half_rows = p.block_n / 2;

  // tile -> this CTA's 128-row slice: (group g, 128-row tile t inside the group) ; t may fall past the group end
  auto my_slice = [&](int pt, int& g, int& t) {
    g = pt / p.tiles_per_group_pairs;
    t = 2 * (pt - g * p.tiles_per_group_pairs) + (int)rank;
  };

  if (warp == 0) {
    // =========================== TMA producer (both CTAs) ===========================
    const bool conv_par = (p.a_mode == SVDX_A_CONV2D) && !p.geglu && !p.wtiles && (BLOCK_M / p.W) <= 32;
    if (conv_par) {
      // warp-wide conv producer (see conv_tile_boxes): lane 0 owns the barriers and the B half-tile
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        const int nt = tile % p.n_tiles;
        const int pt = tile / p.n_tiles;
        int g, t;
        my_slice(pt, g, t);
        const int n0 = nt * bn_out;
        ConvBox bx;
        conv_tile_boxes(p, t, lane, bx);
        const CUtensorMap* amap = bx.lg == 0 ? &p.tma : &p.tma_bh[bx.lg - 1];
        int tap = 0, kci = 0;
        for (int kb = 0; kb < p.kb_total; ++kb) {
          const uint32_t full = bar_full + 8 * stage;
          if (lane == 0) {
            mbar_wait(bar_empty + 8 * stage, phase ^ 1);
            if (leader) mbar_expect_tx(full, 2 * A_STAGE_BYTES + p.block_n * BLOCK_K * 2);
          }
          __syncwarp();
          const int kc = kci * BLOCK_K;
          if (bx.active)
            tma2_load_4d(amap, full, sA + stage * A_STAGE_BYTES + bx.dst_off, kc, p.tap_d0[tap], bx.hh + p.tap_d1[tap],
                         bx.nvalid ? bx.n + p.tap_d2[tap] : (1 << 28));
          if (lane == 0) {
            if constexpr (WIDE) {
              tma2_load_2d(&p.tmb, full, sB + stage * B2_STAGE_BYTES, tap * p.K + kc, n0 + (int)rank * 80);
              tma2_load_2d(&p.tmb, full, sB + stage * B2_STAGE_BYTES + 80 * 128, tap * p.K + kc, n0 + 160 + (int)rank * 80);
            } else {
              tma2_load_2d(&p.tmb, full, sB + stage * B2_STAGE_BYTES, tap * p.K + kc, n0 + (int)rank * b_half_rows);
            }
          }
          if (++kci == p.kb_per_tap) { kci = 0; ++tap; }
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
      }
    } else if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        const int nt = tile % p.n_tiles;
        const int pt = tile / p.n_tiles;
        int g, t;
        my_slice(pt, g, t);
        const int n0 = nt * bn_out;
        int tap = 0, kci = -1;
        for (int kb = 0; kb < p.kb_total; ++kb) {
          if (++kci == p.kb_per_tap) { kci = 0; ++tap; }   // no per-k-block division on the issue path
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          const uint32_t full = bar_full + 8 * stage;
          if (leader) mbar_expect_tx(full, 2 * A_STAGE_BYTES + p.block_n * BLOCK_K * 2);
          const uint32_t dA = sA + stage * A_STAGE_BYTES;
          const uint32_t dB = sB + stage * B2_STAGE_BYTES;
          const int kc = kci * BLOCK_K;
          if (p.a_mode == SVDX_A_ROWS) {
            // rows past the end of the group (t >= tiles_per_group) are out of bounds -> zero
            tma2_load_3d(&p.tma, full, dA, kc, t * BLOCK_M + p.tap_d0[tap], g);
          } else if (p.wtiles) {
            // wide images: 128 consecutive pixels of one image row (see tapgemm.cu)
            const int row = t / p.wtiles;
            const int w0 = (t - row * p.wtiles) * BLOCK_M;
            const int n = row / p.H;
            const int h = row - n * p.H;
            tma2_load_4d(&p.tma, full, dA, kc, w0 + p.tap_d0[tap], h + p.tap_d1[tap], (n < p.nimg) ? n + p.tap_d2[tap] : (1 << 28));
          } else {
            const int R = BLOCK_M / p.W;
            const int dw = p.tap_d0[tap], dh = p.tap_d1[tap], dn = p.tap_d2[tap];
            int rowid = t * R, left = R;   // conv: one group, t indexes 128-pixel tiles directly
            uint32_t dst = dA;
            while (left > 0) {
              const int n = rowid / p.H;
              const int h = rowid - n * p.H;
              const int nn = (n < p.nimg) ? n + dn : (1 << 28);
              int run = min(left, p.H - h);
              int hh = h;
              while (run > 0) {
                int lg = min(p.max_bh_log2, 31 - __clz(run));
                const int bh = 1 << lg;
                tma2_load_4d(lg == 0 ? &p.tma : &p.tma_bh[lg - 1], full, dst, kc, dw, hh + dh, nn);
                dst += bh * p.W * 128;
                hh += bh; run -= bh; left -= bh; rowid += bh;
              }
            }
          }
          // this CTA's half of the B tile: rows [rank*bn/2, +bn/2) of the tile. For GEGLU the tile is
          // [value rows | gate rows], so rank 0 fetches the value rows and rank 1 the gate rows.
          if constexpr (WIDE) {
            // MMA 0 computes columns [n0, n0 + 160) from rows [0, 80) of both CTAs' halves, MMA 1 columns [n0 + 160, n0 + 320)
            tma2_load_2d(&p.tmb, full, dB, tap * p.K + kc, n0 + (int)rank * 80);
            tma2_load_2d(&p.tmb, full, dB + 80 * 128, tap * p.K + kc, n0 + 160 + (int)rank * 80);
          } else if (p.geglu) tma2_load_2d(&p.tmb, full, dB, tap * p.K + kc, rank == 0 ? n0 : p.N / 2 + n0);
          else tma2_load_2d(&p.tmb, full, dB, tap * p.K + kc, n0 + (int)rank * b_half_rows);
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (leader CTA only) ===========================
    if (leader && lane == 0) {
      const uint32_t idesc = make_idesc_bf16(2 * BLOCK_M, WIDE ? 160 : p.block_n, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      uint32_t wide_phase = 0;
      for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
        if constexpr (WIDE) {
          // ONE release barrier: the epilogue of the PREVIOUS tile arrives once the columns both accumulators share are
          // drained (by then the same warps have fully drained the tile before it, which used this tile's columns)
          mbar_wait(bar_tempty, wide_phase ^ 1);
          wide_phase ^= 1;
        } else {
          mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);
        }
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (WIDE ? acc * WIDE_ACC1 : acc * 256);
        for (int kb = 0; kb < p.kb_total; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          const uint32_t aaddr = sA + stage * A_STAGE_BYTES;
          const uint32_t baddr = sB + stage * B2_STAGE_BYTES;
          const uint64_t ad0 = make_smem_desc_sw128(aaddr, 16, 1024), bd0 = make_smem_desc_sw128(baddr, 16, 1024);
          if constexpr (WIDE) {
            const uint64_t bd1 = make_smem_desc_sw128(baddr + 80 * 128, 16, 1024);
#pragma unroll
            for (int j = 0; j < BLOCK_K / 16; ++j) {
              umma2_bf16(d_tmem, ad0 + 2 * j, bd0 + 2 * j, idesc, (kb > 0 || j > 0) ? 1u : 0u);
              umma2_bf16(d_tmem + 160, ad0 + 2 * j, bd1 + 2 * j, idesc, (kb > 0 || j > 0) ? 1u : 0u);
            }
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_K / 16; ++j)   // a 16-deep k-step = +32 bytes = +2 in the start-address field
              umma2_bf16(d_tmem, ad0 + 2 * j, bd0 + 2 * j, idesc, (kb > 0 || j > 0) ? 1u : 0u);
          }
          umma2_commit_mc(bar_empty + 8 * stage, 3);  // frees this smem stage in both CTAs
          if (++stage == STAGES2) { stage = 0; phase ^= 1; }
        }
        umma2_commit_mc(bar_tfull + 8 * acc, 3);  // accumulator complete: wake both epilogues
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else {
    // =========================== epilogue warps (both CTAs, own 128 TMEM lanes) ===========================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    int acc = 0;
    uint32_t acc_phase = 0;
    float s_acc = 1.f, s_r1 = 1.f, s_r2 = 1.f;
    if (p.scales) { s_acc = p.scales[0]; s_r1 = p.scales[1]; s_r2 = p.scales[2]; }
    const int n_out_total = p.geglu ? p.N / 2 : p.N;
    EpiStage st;
    st.base = sEpi + (warp - 2) * EPI_STAGE_BYTES;
    st.off = 0;
    for (int tile = cluster_id; tile < total_tiles; tile += num_clusters) {
      const int nt = tile % p.n_tiles;
      const int pt = tile / p.n_tiles;
      int g, t;
      my_slice(pt, g, t);
      st.grp = g;
      st.row0 = t * BLOCK_M + q * 32;
      const int rin = t * BLOCK_M + q * 32 + lane;
      const bool row_ok = rin < p.rows_per_group;
      const long long m = (long long)g * p.rows_per_group + rin;
      const long long m0 = (long long)g * p.rows_per_group + st.row0;        // fused GroupNorm statistics: first row / rows that exist
      const int valid_rows = max(0, min(32, p.rows_per_group - st.row0));
      if constexpr (EPI == EPI_RES || EPI == EPI_RES_GN || EPI == EPI_FAST_GNB)
        prefetch_epilogue_operands(p, EPI, m0, valid_rows, nt * bn_out, bn_out, n_out_total, lane, half);
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      const uint32_t t_base = tmem_base + (WIDE ? acc * WIDE_ACC1 : acc * 256) + ((uint32_t)(q * 32) << 16);
      if constexpr (WIDE) {
        // accumulator columns shared with the next tile's accumulator first: [192, 320) of an even tile, [0, 128) of an odd one
        const int lo1 = acc ? 0 : WIDE_ACC1, hi1 = acc ? 320 - WIDE_ACC1 : 320;
        const int lo2 = acc ? 320 - WIDE_ACC1 : 0, hi2 = acc ? 320 : WIDE_ACC1;
        if constexpr (EPI == EPI_FAST || EPI == EPI_FAST_GN) epilogue_fast<EPI == EPI_FAST_GN>(p, t_base, m, row_ok, nt * bn_out, half, lo1, hi1, n_out_total, st.base, st.row0, st.grp, lane, m0, valid_rows);
        else if constexpr (EPI == EPI_FAST_GNB) epilogue_fast_gnb(p, t_base, m, row_ok, nt * bn_out, half, lo1, hi1, n_out_total, st.base, st.row0, st.grp, lane, m0, valid_rows);
        else epilogue_res<EPI == EPI_RES_GN>(p, t_base, m, row_ok, nt * bn_out, half, lo1, hi1, n_out_total, s_acc, s_r1, s_r2, st.base, st.off, st.row0, st.grp, lane, m0, valid_rows);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(bar_tempty, 0);
        if constexpr (EPI == EPI_FAST || EPI == EPI_FAST_GN) epilogue_fast<EPI == EPI_FAST_GN>(p, t_base, m, row_ok, nt * bn_out, half, lo2, hi2, n_out_total, st.base, st.row0, st.grp, lane, m0, valid_rows);
        else if constexpr (EPI == EPI_FAST_GNB) epilogue_fast_gnb(p, t_base, m, row_ok, nt * bn_out, half, lo2, hi2, n_out_total, st.base, st.row0, st.grp, lane, m0, valid_rows);
        else epilogue_res<EPI == EPI_RES_GN>(p, t_base, m, row_ok, nt * bn_out, half, lo2, hi2, n_out_total, s_acc, s_r1, s_r2, st.base, st.off, st.row0, st.grp, lane, m0, valid_rows);
        tc_fence_before();
        if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
        continue;
      }
      if constexpr (EPI == EPI_FAST || EPI == EPI_FAST_GN) epilogue_fast<EPI == EPI_FAST_GN>(p, t_base, m, row_ok, nt * bn_out, half, 0, bn_out, n_out_total, st.base, st.row0, st.grp, lane, m0, valid_rows);
      else if constexpr (EPI == EPI_RES || EPI == EPI_RES_GN) epilogue_res<EPI == EPI_RES_GN>(p, t_base, m, row_ok, nt * bn_out, half, 0, bn_out, n_out_total, s_acc, s_r1, s_r2, st.base, st.off, st.row0, st.grp, lane, m0, valid_rows);
      else if constexpr (EPI == EPI_GEGLU) epilogue_geglu(p, t_base, nt * bn_out, half, bn_out, st.base, st.row0, st.grp, lane);
      else if constexpr (EPI == EPI_FAST_GNB) epilogue_fast_gnb(p, t_base, m, row_ok, nt * bn_out, half, 0, bn_out, n_out_total, st.base, st.row0, st.grp, lane, m0, valid_rows);
      else epilogue_tile(p, t_base, m, row_ok, nt * bn_out, half, bn_out, n_out_total, s_acc, s_r1, s_r2, st, lane);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(bar_tempty + 8 * acc, 0);  // the leader's barrier
      if (++acc == ACC_STAGES) { acc = 0; acc_phase ^= 1; }
    }
    if (lane == 0) bulk_wait<0>();
  }

  tc_fence_before();
  cluster_sync_all();   // the peer may still be signalling our barriers / reading our smem until here
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc2(tmem_base, TMEM_COLS);
  }
}

}  // namespace svdx

using namespace svdx;

static int pair_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("SVDX_2CTA");
    mode = (e && e[0] == '0') ? 0 : (e && e[0] == '1') ? 1 : SVDX_2CTA_DEFAULT;
  }
  return mode;
}

bool svdx_tapgemm2_eligible(const SvdxTapGemm* d) {
  if (!pair_mode() || !d) return false;
  if (d->a_major_mn || d->b_major_mn || d->b_mode != 0 || d->split_k != 1) return false;
  if (d->block_n == 320) return !d->geglu && d->N % 320 == 0 && d->M >= 512 && d->out_dtype == SVDX_OUT_BF16;   // WIDE tile
  if (d->block_n < 64 || d->block_n > 256 || d->block_n % 32 || (d->block_n / 2) % 8) return false;
  if (d->geglu && d->block_n % 64) return false;
  if (d->M < 512) return false;                       // small problems: keep the finer 128-row tiling
  const int n_out = d->geglu ? d->N / 2 : d->N;
  const int bn_out = d->geglu ? d->block_n / 2 : d->block_n;
  if (n_out % bn_out) return false;                   // each CTA fetches exactly half a B tile
  return true;
}

int svdx_tapgemm2_launch(const TapGemmKParams& p, cudaStream_t stream) {
  static bool attr_done[SVDX_MAX_DEVICES] = {false};
  const int slot = svdx_device_slot();
  if (!attr_done[slot]) {
    cudaError_t e = cudaFuncSetAttribute(tapgemm2_kernel<EPI_GENERIC>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tapgemm2_kernel<EPI_FAST>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tapgemm2_kernel<EPI_GEGLU>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tapgemm2_kernel<EPI_RES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tapgemm2_kernel<EPI_FAST_GN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tapgemm2_kernel<EPI_RES_GN>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tapgemm2_kernel<EPI_FAST_GNB>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tapgemm2_kernel<EPI_FAST_GNB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2W_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tapgemm2_kernel<EPI_FAST, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2W_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tapgemm2_kernel<EPI_RES, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2W_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tapgemm2_kernel<EPI_FAST_GN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2W_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(tapgemm2_kernel<EPI_RES_GN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM2W_BYTES);
    if (e != cudaSuccess) return svdx_fail_cuda(e, "tapgemm2: set smem attribute");
    attr_done[slot] = true;
  }
  const int total_tiles = p.pair_m_tiles * p.n_tiles;
  int clusters = svdx_num_sms() / 2;
  if (clusters > total_tiles) clusters = total_tiles;
  if (p.block_n == 320) {
    if (p.epi_mode == EPI_FAST_GNB) tapgemm2_kernel<EPI_FAST_GNB, true><<<2 * clusters, NUM_THREADS, SMEM2W_BYTES, stream>>>(p);
    else if (p.epi_mode == EPI_FAST && p.gn_sum) tapgemm2_kernel<EPI_FAST_GN, true><<<2 * clusters, NUM_THREADS, SMEM2W_BYTES, stream>>>(p);
    else if (p.epi_mode == EPI_RES && p.gn_sum) tapgemm2_kernel<EPI_RES_GN, true><<<2 * clusters, NUM_THREADS, SMEM2W_BYTES, stream>>>(p);
    else if (p.epi_mode == EPI_FAST) tapgemm2_kernel<EPI_FAST, true><<<2 * clusters, NUM_THREADS, SMEM2W_BYTES, stream>>>(p);
    else if (p.epi_mode == EPI_RES) tapgemm2_kernel<EPI_RES, true><<<2 * clusters, NUM_THREADS, SMEM2W_BYTES, stream>>>(p);
    else return svdx_fail(SVDX_E_BADARG, "tapgemm2: block_n 320 needs the bf16 TMA-store epilogues");
    cudaError_t e2 = cudaGetLastError();
    if (e2 != cudaSuccess) return svdx_fail_cuda(e2, "tapgemm2: launch (wide)");
    return SVDX_OK;
  }
  if (p.epi_mode == EPI_FAST_GNB) tapgemm2_kernel<EPI_FAST_GNB><<<2 * clusters, NUM_THREADS, SMEM2_BYTES, stream>>>(p);
  else if (p.epi_mode == EPI_FAST && p.gn_sum) tapgemm2_kernel<EPI_FAST_GN><<<2 * clusters, NUM_THREADS, SMEM2_BYTES, stream>>>(p);
  else if (p.epi_mode == EPI_RES && p.gn_sum) tapgemm2_kernel<EPI_RES_GN><<<2 * clusters, NUM_THREADS, SMEM2_BYTES, stream>>>(p);
  else if (p.epi_mode == EPI_FAST) tapgemm2_kernel<EPI_FAST><<<2 * clusters, NUM_THREADS, SMEM2_BYTES, stream>>>(p);
  else if (p.epi_mode == EPI_GEGLU) tapgemm2_kernel<EPI_GEGLU><<<2 * clusters, NUM_THREADS, SMEM2_BYTES, stream>>>(p);
  else if (p.epi_mode == EPI_RES) tapgemm2_kernel<EPI_RES><<<2 * clusters, NUM_THREADS, SMEM2_BYTES, stream>>>(p);
  else tapgemm2_kernel<EPI_GENERIC><<<2 * clusters, NUM_THREADS, SMEM2_BYTES, stream>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return svdx_fail_cuda(e, "tapgemm2: launch");
  return SVDX_OK;
}

// Halo-reuse 3x3 / stride-1 / pad-1 convolution for 128 -> 128 channels (ResNet layer2, resnet.py:39-55) on
// zero-bordered ("padded flat") NHWC fp16 activations X[B][H+2][W+2][128] viewed as a matrix [P][128],
// P = B*(H+2)*(W+2) -- the layer1 recipe (tc_flat3x3.cuh) carried to the layer whose implicit-GEMM launches sat at the
// L2 -> SM cap (nine tap tiles + a pair-split weight tile per k-block = 432 KB of operand traffic per 128 output pixels):
//     Y[p] = act(bias + sum_{kh,kw} W[kh][kw] * X[p + (kh-1)*(W+2) + (kw-1)] (+ R[p]))
//   * one GROUP = two consecutive tiles of 128 flat pixels. Per group and 64-channel half of Cin ONE window of
//     256 + 2*(W+2) + 2 input rows (41 KB) serves all nine taps of both tiles through row-shifted UMMA descriptors;
//   * the weights (nine taps x two Cin halves x [128 cout][64 cin] = 288 KB) cannot stay resident: they stream through a
//     ring of (tap, Cin half) tiles, and every weight tile is used by BOTH tiles of the group before it is released, which
//     halves the weight traffic per output pixel: 2 x 41 KB + 18 x 16 KB = 370 KB per 256 pixels = 185 KB per 128 (was
//     432 KB), of which each CTA of a pair pulls half of the weight part;
//   * CTAs run as PAIRS (tcgen05.mma.cta_group::2, M = 256): the pair's two CTAs hold one tile each in the upper / lower
//     128 accumulator rows and each CTA keeps only HALF of every weight tile (64 cout rows, 8 KB) in its ring. A
//     128x128x16 MMA reads 4 KB of A + 4 KB of B per 64 cycles = 128 B/clk, the whole shared-memory bandwidth of an SM -
//     measured without pairs: tensor pipe 64 % with the TMA writes and the epilogue staging competing for the same
//     banks; the pair form reads 4 + 2 KB per CTA and MMA (96 B/clk), halves the weight bytes each CTA pulls from L2 and
//     deepens the ring (9 half tiles of 8 KB in flight: ~4.6 k MMA cycles of look-ahead). Measured: tensor pipe 87 %;
//   * four TMEM accumulators (2 tiles x 2 groups in flight, all 512 columns), eight epilogue warps, two alternating
//     128B-swizzled staging tiles: the residual of the BasicBlock (resnet.py:52) lands IN the staging tile by TMA one tile
//     ahead and is overwritten in place by bias + residual -> ReLU -> saturating fp16 pack; border pixels are forced back
//     to zero; TMA store.
#pragma once
#include "tc_flat3x3.cuh"

namespace cadre {

struct Halo128Params {
  CUtensorMap tmXa;   // [P][128] box {64, HALO_WIN_A}
  CUtensorMap tmXb;   // [P][128] box {64, 128}
  CUtensorMap tmW;    // [128][1152] box {64, 64}: one CTA's half (64 cout rows) of a (tap, Cin half) weight tile
  CUtensorMap tmY;    // [P][128] box {64, 128}
  CUtensorMap tmR;    // residual [P][128] box {64, 128} (valid when res != nullptr)
  int P, H, W, PW;    // PW = W + 2
  int num_tiles, num_groups, num_pairs;   // a cluster of two CTAs takes the group pair (2m, 2m + 1)
  const float* bias;
  const enc_t* res;   // padded flat, 128 channels, or nullptr
  int act;
};

constexpr int HALO_WIN_A = 200;                               // rows of the first TMA box (25 KB: keeps the second 1 KB aligned)
constexpr int HALO_WIN_ROWS = HALO_WIN_A + 128;               // 328 >= 256 + 2*35 + 2
constexpr int HALO_WIN_BYTES = HALO_WIN_ROWS * 128;           // one Cin half of a group's window
constexpr int HALO_W_BYTES = 64 * 128;                        // this CTA's half (64 cout rows) of a (tap, Cin half) weight tile
constexpr int HALO_NW = 9;                                    // weight tiles in flight (what fits next to the residual tile)
constexpr int HALO_OUT_BYTES = 2 * 128 * 128;                 // one output tile: two 64-channel groups
constexpr int HALO_RES_BYTES = HALO_OUT_BYTES;                // second staging tile (tiles alternate between the two)
constexpr int HALO_NBAR = 2 + 2 + 2 * HALO_NW + 4 + 4 + 2;
constexpr int HALO_SMEM =
    2 * HALO_WIN_BYTES + HALO_NW * HALO_W_BYTES + HALO_OUT_BYTES + HALO_RES_BYTES + HALO_NBAR * 8 + 16 + 1024;

__global__ void __launch_bounds__(320, 1) tc_halo128_kernel(const __grid_constant__ Halo128Params p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* win_s = smem;                                   // 2 x 41 KB (slot = Cin half)
  uint8_t* w_s = win_s + 2 * HALO_WIN_BYTES;               // HALO_NW x 16 KB
  uint8_t* out_s = w_s + HALO_NW * HALO_W_BYTES;           // 32 KB
  uint8_t* res_s = out_s + HALO_OUT_BYTES;                 // 32 KB: staging tile of the odd tiles
  uint64_t* win_full = reinterpret_cast<uint64_t*>(res_s + HALO_RES_BYTES);   // [2]
  uint64_t* win_empty = win_full + 2;                      // [2]
  uint64_t* w_full = win_empty + 2;                        // [HALO_NW]
  uint64_t* w_empty = w_full + HALO_NW;                    // [HALO_NW]
  uint64_t* tfull = w_empty + HALO_NW;                     // [4] accumulator = (group parity) * 2 + tile of the group
  uint64_t* tempty = tfull + 4;                            // [4]
  uint64_t* res_full = tempty + 4;                         // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + 2);

  __shared__ __align__(16) float s_bias[128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 128) s_bias[threadIdx.x] = p.bias[threadIdx.x];
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmXa);
    tma_prefetch_desc(&p.tmXb);
    tma_prefetch_desc(&p.tmW);
    tma_prefetch_desc(&p.tmY);
    if (p.res != nullptr) tma_prefetch_desc(&p.tmR);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&win_full[i], 1);
      mbar_init(&win_empty[i], 1);
    }
    for (int i = 0; i < HALO_NW; ++i) {
      mbar_init(&w_full[i], 1);    // used in the leader CTA only: both CTAs' halves complete bytes on it
      mbar_init(&w_empty[i], 1);   // the leader's commit arrives on both CTAs' barriers
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 16);  // leader CTA: one arrival per epilogue warp of both CTAs
    }
    mbar_init(&res_full[0], 1);
    mbar_init(&res_full[1], 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, 512);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();   // both CTAs' barriers exist before either signals the other
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int crank = static_cast<int>(cluster_ctarank());
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (converged warp, elected lane):
    // windows and weight tiles in exactly the order the MMA warp consumes them
    pdl_wait();   // the activations are written by the stream predecessor
    int lg = 0;
    unsigned wq = 0;   // weight tiles issued
    for (int m = cluster_id; m < p.num_pairs; m += num_clusters, ++lg) {
      const int g = 2 * m + crank;           // g >= num_groups (odd group count): a dummy group of zero-filled windows
      const int row0 = g * 256 - p.PW - 1;   // may be negative / run past P: TMA zero-fills out-of-range rows
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        mbar_wait(&win_empty[h], (lg & 1) ^ 1);
        if (elect_one()) {
          // both CTAs load into their own shared memory; all bytes are accounted on the LEADER's barrier
          if (crank == 0) mbar_expect_tx(&win_full[h], 2 * HALO_WIN_BYTES);
          tma_load_2d_2sm(win_s + h * HALO_WIN_BYTES, &p.tmXa, &win_full[h], h * 64, row0);
          tma_load_2d_2sm(win_s + h * HALO_WIN_BYTES + HALO_WIN_A * 128, &p.tmXb, &win_full[h], h * 64, row0 + HALO_WIN_A);
        }
        __syncwarp();
#pragma unroll 1
        for (int t = 0; t < 9; ++t, ++wq) {
          const int slot = wq % HALO_NW;
          mbar_wait(&w_empty[slot], ((wq / HALO_NW) & 1) ^ 1);
          if (elect_one()) {
            if (crank == 0) mbar_expect_tx(&w_full[slot], 2 * HALO_W_BYTES);
            tma_load_2d_2sm(w_s + slot * HALO_W_BYTES, &p.tmW, &w_full[slot], t * 128 + h * 64, crank * 64);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    if (crank == 0) {
    // ------------------------------------------------------------ MMA issuer (leader CTA of the pair; converged warp,
    // elected lane). The descriptors address the same offsets in both CTAs: rows 0..127 of every M = 256 MMA are the
    // leader's tile, rows 128..255 the peer's, each against its own window; B rows 0..63 / 64..127 likewise.
    constexpr uint32_t idesc = umma_idesc(CADRE_ENC_FP16 ? 0u : 1u, 0, 0, 256, 128);
    const uint32_t win_addr0 = smem_u32(win_s), w_addr0 = smem_u32(w_s);
    int lg = 0;
    unsigned wq = 0;
    for (int m = cluster_id; m < p.num_pairs; m += num_clusters, ++lg) {
      const int buf = lg & 1;
      const uint32_t aph = (lg >> 1) & 1;
      mbar_wait(&tempty[buf * 2 + 0], aph ^ 1);
      mbar_wait(&tempty[buf * 2 + 1], aph ^ 1);
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        mbar_wait(&win_full[h], lg & 1);
        const uint32_t win_addr = win_addr0 + h * HALO_WIN_BYTES;
#pragma unroll 1
        for (int t = 0; t < 9; ++t, ++wq) {
          const int slot = wq % HALO_NW;
          mbar_wait(&w_full[slot], (wq / HALO_NW) & 1);
          tc_fence_after();
          if (elect_one()) {
            const int kh = t / 3, kw = t - kh * 3;
            const uint32_t b_addr = w_addr0 + slot * HALO_W_BYTES;
#pragma unroll
            for (int s = 0; s < 2; ++s) {
              const uint32_t a_addr = win_addr + (s * 128 + kh * p.PW + kw) * 128;
              const uint32_t tacc = tmem_base + (buf * 2 + s) * 128;
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const uint64_t da = umma_smem_desc(a_addr + k * 32, 16, 1024, 2);
                const uint64_t db = umma_smem_desc(b_addr + k * 32, 16, 1024, 2);
                tc_mma_f16_2sm(tacc, da, db, idesc, (h | t | k) != 0);
              }
            }
            tc_commit_2sm(&w_empty[slot]);
            if (t == 8) tc_commit_2sm(&win_empty[h]);
            if (t == 8 && h == 1) {
              tc_commit_2sm(&tfull[buf * 2 + 0]);
              tc_commit_2sm(&tfull[buf * 2 + 1]);
            }
          }
          __syncwarp();
        }
      }
    }
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------ epilogue: 8 warps, 2 per TMEM lane quarter
    pdl_wait();
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;     // which 64 of the 128 output channels
    const int row = q * 32 + lane;
    const bool leader = (warp == 2 && lane == 0);
    const int img_pix = (p.H + 2) * p.PW;
    const bool has_res = p.res != nullptr;
    // Two staging tiles, used alternately. The BasicBlock residual (resnet.py:52) of a tile arrives by TMA IN its staging
    // tile, one tile ahead: a thread finds its row's 16-byte chunks exactly where it will write the result, adds and
    // overwrites them in place. (Row-per-thread global loads - 32 cache lines per warp instruction - cost +26 us per conv,
    // one residual buffer requested after the previous tile's epilogue left ~1.5 us exposed on every second tile.)
    auto request_residual = [&](int tile, int b) {
      uint8_t* dst = b ? res_s : out_s;
      mbar_expect_tx(&res_full[b], HALO_RES_BYTES);
      tma_load_2d(dst, &p.tmR, &res_full[b], 0, tile * 128);
      tma_load_2d(dst + 128 * 128, &p.tmR, &res_full[b], 64, tile * 128);
    };
    if (has_res && leader && cluster_id < p.num_pairs) request_residual((2 * cluster_id + crank) * 2, 0);
    int lg = 0, lt = 0;
    for (int m = cluster_id; m < p.num_pairs; m += num_clusters, ++lg) {
      const int g = 2 * m + crank;
      const int buf = lg & 1;
      const uint32_t aph = (lg >> 1) & 1;
#pragma unroll 1
      for (int s = 0; s < 2; ++s, ++lt) {
        const int tile = g * 2 + s;
        const long long pix = static_cast<long long>(tile) * 128 + row;
        const int rem = static_cast<int>(pix % img_pix);
        const int y = rem / p.PW, x = rem - y * p.PW;
        const bool interior = pix < p.P && y >= 1 && y <= p.H && x >= 1 && x <= p.W;
        uint8_t* stage = (lt & 1) ? res_s : out_s;
        if (leader) {
          // this tile's staging buffer was last read by the store of tile lt-2; with a residual the OTHER buffer (store
          // of tile lt-1) must be free as well: the next tile's residual lands there during this tile's epilogue
          if (has_res) tma_store_wait_read(); else tma_store_wait_read1();
          if (has_res) {
            const int next = s == 0 ? tile + 1 : (m + num_clusters < p.num_pairs ? (2 * (m + num_clusters) + crank) * 2 : -1);
            if (next >= 0) request_residual(next, (lt + 1) & 1);
          }
        }
        epi_bar_sync256();
        const int acc = buf * 2 + s;
        mbar_wait(&tfull[acc], aph);
        tc_fence_after();
        if (has_res) mbar_wait(&res_full[lt & 1], (lt >> 1) & 1);
        const uint32_t taddr = tmem_base + acc * 128 + half * 64 + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          if (c == 1) {   // this warp's part of the accumulator is in registers
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tempty[acc]);
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int col = half * 64 + c * 32 + 8 * j;     // first of 8 output channels
            const float4 b_lo = *reinterpret_cast<const float4*>(&s_bias[col]);
            const float4 b_hi = *reinterpret_cast<const float4*>(&s_bias[col + 4]);
            const float bb[8] = {b_lo.x, b_lo.y, b_lo.z, b_lo.w, b_hi.x, b_hi.y, b_hi.z, b_hi.w};
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[8 * j + i]) + bb[i];
            const int chunk = c * 4 + j;                // 16-byte chunk inside this thread's 64-channel group
            const int soff = half * (128 * 128) + row * 128 + ((chunk ^ (row & 7)) << 4);
            if (has_res) {
              const uint4 rr = *reinterpret_cast<const uint4*>(stage + soff);
              const enc_t* h8 = reinterpret_cast<const enc_t*>(&rr);
#pragma unroll
              for (int i = 0; i < 8; ++i) v[i] += enc_to_float(h8[i]);
            }
            uint4 u;
            if (p.act == ACT_RELU) {
              u.x = enc_pack2_relu(v[0], v[1]), u.y = enc_pack2_relu(v[2], v[3]);
              u.z = enc_pack2_relu(v[4], v[5]), u.w = enc_pack2_relu(v[6], v[7]);
            } else {
              u.x = enc_pack2(v[0], v[1]), u.y = enc_pack2(v[2], v[3]);
              u.z = enc_pack2(v[4], v[5]), u.w = enc_pack2(v[6], v[7]);
            }
            if (!interior) u = make_uint4(0, 0, 0, 0);  // keep the zero border intact
            *reinterpret_cast<uint4*>(stage + soff) = u;
          }
        }
        fence_proxy_async_smem();
        epi_bar_sync256();
        if (leader && tile < p.num_tiles) {   // an odd tile count leaves the last group's second tile empty
          tma_store_2d(&p.tmY, stage, 0, tile * 128);
          tma_store_2d(&p.tmY, stage + 128 * 128, 64, tile * 128);
          tma_store_commit();
        }
      }
    }
    if (leader) tma_store_wait_all();
  }

  tc_fence_before();
  cluster_sync_all();   // the peer may still arrive on this CTA's barriers / the leader's MMAs read the peer's smem
  if (warp == 1) tmem_dealloc_2sm(tmem_base, 512);
}

}  // namespace cadre

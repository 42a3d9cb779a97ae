// Halo-reuse 3x3 / stride-1 / pad-1 convolution for 128 -> 128 channels (ResNet layer2, resnet.py:39-55) on
// zero-bordered ("padded flat") NHWC fp16 activations X[B][H+2][W+2][128] viewed as a matrix [P][128],
// P = B*(H+2)*(W+2) -- the layer1 recipe (tc_flat3x3.cuh) carried to the layer whose implicit-GEMM launches sat at the
// L2 -> SM cap (nine tap tiles + a pair-split weight tile per k-block = 432 KB of operand traffic per 128 output pixels):
//     Y[p] = act(bias + sum_{kh,kw} W[kh][kw] * X[p + (kh-1)*(W+2) + (kw-1)] (+ R[p]))
//   * one GROUP = two consecutive tiles of 128 flat pixels. Per group and 64-channel half of Cin ONE window of
//     256 + 2*(W+2) + 2 input rows (41 KB) serves all nine taps of both tiles through row-shifted UMMA descriptors;
//   * the weights (nine taps x two Cin halves x [128 cout][64 cin] = 288 KB) cannot stay resident: they stream through a
//     ring of 16 KB tiles, and every weight tile is used by BOTH tiles of the group before it is released, which halves
//     the weight traffic per output pixel: 2 x 41 KB + 18 x 16 KB = 370 KB per 256 pixels = 185 KB per 128 (was 432 KB);
//   * four TMEM accumulators (2 tiles x 2 groups in flight, all 512 columns), eight epilogue warps: the residual of the
//     BasicBlock is prefetched from global memory while the MMAs run (the main loop is 4x longer per tile than layer1's,
//     the epilogue has the slack), bias / ReLU / saturating fp16 pack, border pixels forced back to zero, 128B-swizzled
//     staging tile, TMA store.
#pragma once
#include "tc_flat3x3.cuh"

namespace cadre {

struct Halo128Params {
  CUtensorMap tmXa;   // [P][128] box {64, HALO_WIN_A}
  CUtensorMap tmXb;   // [P][128] box {64, 128}
  CUtensorMap tmW;    // [128][1152] box {64, 128}
  CUtensorMap tmY;    // [P][128] box {64, 128}
  int P, H, W, PW;    // PW = W + 2
  int num_tiles, num_groups;
  const float* bias;
  const enc_t* res;   // padded flat, 128 channels, or nullptr
  int act;
};

constexpr int HALO_WIN_A = 200;                               // rows of the first TMA box (25 KB: keeps the second 1 KB aligned)
constexpr int HALO_WIN_ROWS = HALO_WIN_A + 128;               // 328 >= 256 + 2*35 + 2
constexpr int HALO_WIN_BYTES = HALO_WIN_ROWS * 128;           // one Cin half of a group's window
constexpr int HALO_W_BYTES = 128 * 128;                       // one (tap, Cin half) weight tile
constexpr int HALO_NW = 6;                                    // weight tiles in flight
constexpr int HALO_OUT_BYTES = 2 * 128 * 128;                 // one output tile: two 64-channel groups
constexpr int HALO_NBAR = 2 + 2 + 2 * HALO_NW + 4 + 4;
constexpr int HALO_SMEM = 2 * HALO_WIN_BYTES + HALO_NW * HALO_W_BYTES + HALO_OUT_BYTES + HALO_NBAR * 8 + 16 + 1024;

__global__ void __launch_bounds__(320, 1) tc_halo128_kernel(const __grid_constant__ Halo128Params p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* win_s = smem;                                   // 2 x 41 KB (slot = Cin half)
  uint8_t* w_s = win_s + 2 * HALO_WIN_BYTES;               // HALO_NW x 16 KB
  uint8_t* out_s = w_s + HALO_NW * HALO_W_BYTES;           // 32 KB
  uint64_t* win_full = reinterpret_cast<uint64_t*>(out_s + HALO_OUT_BYTES);   // [2]
  uint64_t* win_empty = win_full + 2;                      // [2]
  uint64_t* w_full = win_empty + 2;                        // [HALO_NW]
  uint64_t* w_empty = w_full + HALO_NW;                    // [HALO_NW]
  uint64_t* tfull = w_empty + HALO_NW;                     // [4] accumulator = (group parity) * 2 + tile of the group
  uint64_t* tempty = tfull + 4;                            // [4]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 4);

  __shared__ __align__(16) float s_bias[128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 128) s_bias[threadIdx.x] = p.bias[threadIdx.x];
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmXa);
    tma_prefetch_desc(&p.tmXb);
    tma_prefetch_desc(&p.tmW);
    tma_prefetch_desc(&p.tmY);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&win_full[i], 1);
      mbar_init(&win_empty[i], 1);
    }
    for (int i = 0; i < HALO_NW; ++i) {
      mbar_init(&w_full[i], 1);
      mbar_init(&w_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);   // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (converged warp, elected lane):
    // windows and weight tiles in exactly the order the MMA warp consumes them
    pdl_wait();   // the activations are written by the stream predecessor
    int lg = 0;
    unsigned wq = 0;   // weight tiles issued
    for (int g = blockIdx.x; g < p.num_groups; g += gridDim.x, ++lg) {
      const int row0 = g * 256 - p.PW - 1;   // may be negative / run past P: TMA zero-fills out-of-range rows
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        mbar_wait(&win_empty[h], (lg & 1) ^ 1);
        if (elect_one()) {
          mbar_expect_tx(&win_full[h], HALO_WIN_BYTES);
          tma_load_2d(win_s + h * HALO_WIN_BYTES, &p.tmXa, &win_full[h], h * 64, row0);
          tma_load_2d(win_s + h * HALO_WIN_BYTES + HALO_WIN_A * 128, &p.tmXb, &win_full[h], h * 64, row0 + HALO_WIN_A);
        }
        __syncwarp();
#pragma unroll 1
        for (int t = 0; t < 9; ++t, ++wq) {
          const int slot = wq % HALO_NW;
          mbar_wait(&w_empty[slot], ((wq / HALO_NW) & 1) ^ 1);
          if (elect_one()) {
            mbar_expect_tx(&w_full[slot], HALO_W_BYTES);
            tma_load_2d(w_s + slot * HALO_W_BYTES, &p.tmW, &w_full[slot], t * 128 + h * 64, 0);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (converged warp, elected lane)
    constexpr uint32_t idesc = umma_idesc(CADRE_ENC_FP16 ? 0u : 1u, 0, 0, 128, 128);
    const uint32_t win_addr0 = smem_u32(win_s), w_addr0 = smem_u32(w_s);
    int lg = 0;
    unsigned wq = 0;
    for (int g = blockIdx.x; g < p.num_groups; g += gridDim.x, ++lg) {
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
                tc_mma_f16(tacc, da, db, idesc, (h | t | k) != 0);
              }
            }
            tc_commit(&w_empty[slot]);
            if (t == 8) tc_commit(&win_empty[h]);
            if (t == 8 && h == 1) {
              tc_commit(&tfull[buf * 2 + 0]);
              tc_commit(&tfull[buf * 2 + 1]);
            }
          }
          __syncwarp();
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
    int lg = 0;
    for (int g = blockIdx.x; g < p.num_groups; g += gridDim.x, ++lg) {
      const int buf = lg & 1;
      const uint32_t aph = (lg >> 1) & 1;
#pragma unroll 1
      for (int s = 0; s < 2; ++s) {
        const int tile = g * 2 + s;
        const long long pix = static_cast<long long>(tile) * 128 + row;
        const int rem = static_cast<int>(pix % img_pix);
        const int y = rem / p.PW, x = rem - y * p.PW;
        const bool interior = pix < p.P && y >= 1 && y <= p.H && x >= 1 && x <= p.W;
        // this thread's residual half-row (64 channels = 128 bytes), requested before the accumulator is ready
        uint4 rres[8];
        if (has_res) {
          const uint4* rp = reinterpret_cast<const uint4*>(p.res + pix * 128 + half * 64);
#pragma unroll
          for (int j = 0; j < 8; ++j) rres[j] = interior ? __ldg(rp + j) : make_uint4(0, 0, 0, 0);
        }
        if (leader) tma_store_wait_read();   // the previous tile's store has read the staging buffer
        epi_bar_sync256();
        const int acc = buf * 2 + s;
        mbar_wait(&tfull[acc], aph);
        tc_fence_after();
        const uint32_t taddr = tmem_base + acc * 128 + half * 64 + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
          if (c == 1) {   // this warp's part of the accumulator is in registers
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[acc]);
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
            if (has_res) {
              const enc_t* h8 = reinterpret_cast<const enc_t*>(&rres[c * 4 + j]);
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
            const int chunk = c * 4 + j;                // 16-byte chunk inside this thread's 64-channel group
            *reinterpret_cast<uint4*>(out_s + half * (128 * 128) + row * 128 + ((chunk ^ (row & 7)) << 4)) = u;
          }
        }
        fence_proxy_async_smem();
        epi_bar_sync256();
        if (leader && tile < p.num_tiles) {   // an odd tile count leaves the last group's second tile empty
          tma_store_2d(&p.tmY, out_s, 0, tile * 128);
          tma_store_2d(&p.tmY, out_s + 128 * 128, 64, tile * 128);
          tma_store_commit();
        }
      }
    }
    if (leader) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

}  // namespace cadre

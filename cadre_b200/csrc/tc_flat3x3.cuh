// Halo-reuse 3x3 / stride-1 / pad-1 convolution for 64 -> 64 channels (ResNet layer1, resnet.py:39-55) on
// zero-bordered ("padded flat") NHWC fp16 activations X[B][H+2][W+2][64] viewed as a matrix [P][64],
// P = B*(H+2)*(W+2). With the border stored explicitly the convolution is a sum of nine ROW-SHIFTED GEMMs:
//     Y[p] = act(bias + sum_{kh,kw} W[kh][kw] * X[p + (kh-1)*(W+2) + (kw-1)] (+ R[p]))
// for every flat pixel p, border pixels being forced back to zero. One tile = 128 consecutive flat pixels:
//   * ONE TMA window of 128 + 2*(W+2) + 2 input rows per tile serves all nine taps: the A operand of tap
//     (kh,kw) is the same shared-memory window read through a UMMA descriptor whose start address is shifted
//     by (kh*(W+2) + kw) rows of 128 bytes (verified on B200: SWIZZLE_128B descriptors may start at any 128-byte
//     row with base_offset 0, the swizzle is a function of the absolute shared-memory address);
//   * the nine 64x64 weight tiles (72 KB) are loaded once per CTA and stay resident;
//   * persistent CTAs, double-buffered windows and TMEM accumulators, TMA-store epilogue as in tc_persist.cuh.
// Per tile the tensor core runs 36 MMAs (128x64x16) against 34 KB of input traffic instead of 9 x 24 KB.
#pragma once
#include "tc_persist.cuh"

namespace cadre {

struct FlatParams {
  CUtensorMap tmX;    // [P][64] box {64, WIN_A} and {64, 128}: two maps because the box sizes differ
  CUtensorMap tmX2;
  CUtensorMap tmW;    // [64][576] box {64, 64}
  CUtensorMap tmY;    // [P][64] box {64, 128}
  int P, H, W, PW;    // PW = W + 2
  int num_tiles;
  const float* bias;
  const enc_t* res;   // padded flat, 64 channels
  int act;
};

constexpr int FLAT_WIN_A = 136;               // rows of the first TMA box (17 KB, keeps the second box 1 KB aligned)
constexpr int FLAT_WIN_ROWS = FLAT_WIN_A + 128;  // 264 >= 128 + 2*66 + 2
constexpr int FLAT_WIN_BYTES = FLAT_WIN_ROWS * 128;
constexpr int FLAT_W_BYTES = 9 * 64 * 128;
constexpr int FLAT_NWIN = 3;  // input windows in flight (a 33 KB window takes longer to land than 36 MMAs take to run)
constexpr int FLAT_SMEM = FLAT_W_BYTES + FLAT_NWIN * FLAT_WIN_BYTES + 128 * 128 + 16 * 8 + 16 + 1024;

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}

__global__ void __launch_bounds__(320, 1) tc_flat3x3_kernel(const __grid_constant__ FlatParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_s = smem;                                   // 9 x 8 KB
  uint8_t* win_s = w_s + FLAT_W_BYTES;                   // 2 x 33 KB
  uint8_t* out_s = win_s + FLAT_NWIN * FLAT_WIN_BYTES;   // 16 KB
  uint64_t* w_full = reinterpret_cast<uint64_t*>(out_s + 128 * 128);
  uint64_t* win_full = w_full + 1;   // [FLAT_NWIN]
  uint64_t* win_empty = win_full + FLAT_NWIN;
  uint64_t* tfull = win_empty + FLAT_NWIN;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  __shared__ float s_bias[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 64) s_bias[threadIdx.x] = p.bias[threadIdx.x];
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmX);
    tma_prefetch_desc(&p.tmX2);
    tma_prefetch_desc(&p.tmW);
    tma_prefetch_desc(&p.tmY);
    mbar_init(w_full, 1);
    for (int a = 0; a < FLAT_NWIN; ++a) {
      mbar_init(&win_full[a], 1);
      mbar_init(&win_empty[a], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], 8);  // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 128);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------ TMA producer
    mbar_expect_tx(w_full, FLAT_W_BYTES);
    for (int t = 0; t < 9; ++t) tma_load_2d(w_s + t * 8192, &p.tmW, w_full, t * 64, 0);
    int lt = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++lt) {
      const int b = lt % FLAT_NWIN;
      const uint32_t ph = (lt / FLAT_NWIN) & 1;
      mbar_wait(&win_empty[b], ph ^ 1);
      mbar_expect_tx(&win_full[b], FLAT_WIN_BYTES);
      const int row0 = tile * 128 - p.PW - 1;  // may be negative: TMA zero-fills out-of-range rows
      tma_load_2d(win_s + b * FLAT_WIN_BYTES, &p.tmX, &win_full[b], 0, row0);
      tma_load_2d(win_s + b * FLAT_WIN_BYTES + FLAT_WIN_A * 128, &p.tmX2, &win_full[b], 0, row0 + FLAT_WIN_A);
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc(CADRE_ENC_FP16 ? 0u : 1u, 0, 0, 128, 64);
    mbar_wait(w_full, 0);
    const uint32_t w_addr = smem_u32(w_s);
    int lt = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++lt) {
      const int b = lt & 1;
      const uint32_t ph = (lt >> 1) & 1;
      const int wb = lt % FLAT_NWIN;
      mbar_wait(&tempty[b], ph ^ 1);
      mbar_wait(&win_full[wb], (lt / FLAT_NWIN) & 1);
      tc_fence_after();
      const uint32_t win_addr = smem_u32(win_s + wb * FLAT_WIN_BYTES);
      const uint32_t tacc = tmem_base + b * 64;
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const int kh = t / 3, kw = t - kh * 3;
        const uint32_t a_addr = win_addr + (kh * p.PW + kw) * 128;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = umma_smem_desc(a_addr + k * 32, 16, 1024, 2);
          const uint64_t db = umma_smem_desc(w_addr + t * 8192 + k * 32, 16, 1024, 2);
          tc_mma_f16(tacc, da, db, idesc, (t | k) != 0);
        }
      }
      tc_commit(&win_empty[wb]);
      tc_commit(&tfull[b]);
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------ epilogue: 8 warps, 2 per TMEM lane quarter
    // (with 4 warps the ~600 instructions per thread and tile were the bottleneck: 3.3 k cycles per tile
    //  against 1.7 k cycles of MMA; measured on B200)
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;     // which 32 of the 64 output channels
    const int row = q * 32 + lane;
    const bool leader = (warp == 2 && lane == 0);
    const int img_pix = (p.H + 2) * p.PW;
    float bias_r[32];                     // this thread always handles the same 32 channels
#pragma unroll
    for (int i = 0; i < 32; ++i) bias_r[i] = s_bias[half * 32 + i];
    int lt = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++lt) {
      const int b = lt & 1;
      const uint32_t ph = (lt >> 1) & 1;
      const int pix = tile * 128 + row;
      const int rem = pix % img_pix;
      const int y = rem / p.PW, x = rem - y * p.PW;
      const bool interior = pix < p.P && y >= 1 && y <= p.H && x >= 1 && x <= p.W;
      uint4 rres[4];
      const bool has_res = p.res != nullptr;
      if (has_res) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.res + static_cast<long long>(pix) * 64 + half * 32);
#pragma unroll
        for (int j = 0; j < 4; ++j) rres[j] = interior ? __ldg(rp + j) : make_uint4(0, 0, 0, 0);
      }
      if (leader) tma_store_wait_read();
      epi_bar_sync256();
      mbar_wait(&tfull[b], ph);
      tc_fence_after();
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + b * 64 + half * 32 + (static_cast<uint32_t>(q * 32) << 16), r);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[b]);
      uint8_t* rowp = out_s + row * 128;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[8 * j + i]) + bias_r[8 * j + i];
        if (has_res) {
          const enc_t* h8 = reinterpret_cast<const enc_t*>(&rres[j]);
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] += enc_to_float(h8[i]);
        }
        uint4 u;
        if (p.act == ACT_RELU) {
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
          u.x = enc_pack2_pos(v[0], v[1]), u.y = enc_pack2_pos(v[2], v[3]);
          u.z = enc_pack2_pos(v[4], v[5]), u.w = enc_pack2_pos(v[6], v[7]);
        } else {
          u.x = enc_pack2(v[0], v[1]), u.y = enc_pack2(v[2], v[3]);
          u.z = enc_pack2(v[4], v[5]), u.w = enc_pack2(v[6], v[7]);
        }
        if (!interior) u = make_uint4(0, 0, 0, 0);  // keep the zero border intact
        const int chunk = half * 4 + j;
        *reinterpret_cast<uint4*>(rowp + ((chunk ^ (row & 7)) << 4)) = u;
      }
      fence_proxy_async_smem();
      epi_bar_sync256();
      if (leader) {
        tma_store_2d(&p.tmY, out_s, 0, tile * 128);
        tma_store_commit();
      }
    }
    if (leader) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 128);
}

}  // namespace cadre

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
//   * persistent CTAs, double-buffered windows, TMEM accumulators and output staging, TMA-store epilogue;
//   * the residual of a BasicBlock (resnet.py:52) is added ON THE TENSOR CORE: its 128x64 tile rides in the
//     same TMA transaction as the window and is multiplied by a 64x64 identity (a tenth "tap", exact in
//     fp16 x 1.0 -> fp32), because row-per-thread residual loads made the epilogue the critical path
//     (+40 us per conv, measured on B200).
// Per tile the tensor core runs 36 (+4) MMAs (128x64x16) against 34 KB of input traffic instead of 9 x 24 KB.
#pragma once
#include "tc_persist.cuh"

namespace cadre {

struct FlatParams {
  CUtensorMap tmX;    // [P][64] box {64, WIN_A} and {64, 128}: two maps because the box sizes differ
  CUtensorMap tmX2;
  CUtensorMap tmW;    // [64][576] box {64, 64}
  CUtensorMap tmY;    // [P][64] box {64, 128}
  CUtensorMap tmR;    // residual [P][64] box {64, 128} (valid when res != nullptr)
  int P, H, W, PW;    // PW = W + 2
  int num_tiles;
  const float* bias;
  const enc_t* res;   // padded flat, 64 channels
  int act;
  long long* dbg;    // optional per-CTA cycle counters [16] (cadre_debug_clk), nullptr in production
};

constexpr int FLAT_WIN_A = 136;               // rows of the first TMA box (17 KB, keeps the second box 1 KB aligned)
constexpr int FLAT_WIN_ROWS = FLAT_WIN_A + 128;  // 264 >= 128 + 2*66 + 2
constexpr int FLAT_WIN_BYTES = FLAT_WIN_ROWS * 128;
constexpr int FLAT_RES_BYTES = 128 * 128;                    // residual tile, directly behind the window
constexpr int FLAT_SLOT_BYTES = FLAT_WIN_BYTES + FLAT_RES_BYTES;
constexpr int FLAT_W_BYTES = 10 * 64 * 128;                  // nine taps + the identity
constexpr int FLAT_NWIN = 2;  // window (+ residual) slots in flight (three measured no faster on B200)
constexpr int FLAT_SMEM = FLAT_W_BYTES + FLAT_NWIN * FLAT_SLOT_BYTES + 2 * 128 * 128 + 16 * 8 + 16 + 1024;

__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}

__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* m, uint64_t* leader_bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], "
      "[%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}

__global__ void __launch_bounds__(320, 1) tc_flat3x3_kernel(const __grid_constant__ FlatParams p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* w_s = smem;                                   // 9 x 8 KB taps + 8 KB identity
  uint8_t* win_s = w_s + FLAT_W_BYTES;                   // 2 x (33 KB window + 16 KB residual)
  uint8_t* out_s = win_s + FLAT_NWIN * FLAT_SLOT_BYTES;  // 2 x 16 KB
  uint64_t* w_full = reinterpret_cast<uint64_t*>(out_s + 2 * 128 * 128);
  uint64_t* win_full = w_full + 1;   // [FLAT_NWIN]
  uint64_t* win_empty = win_full + FLAT_NWIN;
  uint64_t* tfull = win_empty + FLAT_NWIN;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  __shared__ float s_bias[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool has_res = p.res != nullptr;
  if (threadIdx.x < 64) s_bias[threadIdx.x] = p.bias[threadIdx.x];
  if (has_res) {
    // identity B tile [n = 64][k = 64] in the SWIZZLE_128B K-major layout: row n, 16-byte chunk (k/8) ^ (n&7)
    uint4* id4 = reinterpret_cast<uint4*>(w_s + 9 * 8192);
    for (int i = threadIdx.x; i < 512; i += blockDim.x) id4[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    if (threadIdx.x < 64) {
      const int n = threadIdx.x;
      enc_t* rowp = reinterpret_cast<enc_t*>(w_s + 9 * 8192 + n * 128 + (((n >> 3) ^ (n & 7)) << 4));
      rowp[n & 7] = enc_from_float(1.f);
    }
    fence_proxy_async_smem();
  }
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

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (converged warp, elected lane)
    if (elect_one()) {
      mbar_expect_tx(w_full, 9 * 8192);
      for (int t = 0; t < 9; ++t) tma_load_2d(w_s + t * 8192, &p.tmW, w_full, t * 64, 0);
    }
    __syncwarp();
    pdl_wait();   // weights / bias are never written by a stream predecessor; activations are
    int lt = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++lt) {
      const int b = lt % FLAT_NWIN;
      const uint32_t ph = (lt / FLAT_NWIN) & 1;
      const long long t0 = p.dbg ? clock64() : 0;
      mbar_wait(&win_empty[b], ph ^ 1);
      if (p.dbg && lane == 0) p.dbg[blockIdx.x * 16 + 1] += clock64() - t0;
      if (elect_one()) {
      mbar_expect_tx(&win_full[b], has_res ? FLAT_SLOT_BYTES : FLAT_WIN_BYTES);
      const int row0 = tile * 128 - p.PW - 1;  // may be negative: TMA zero-fills out-of-range rows
      tma_load_2d(win_s + b * FLAT_SLOT_BYTES, &p.tmX, &win_full[b], 0, row0);
      tma_load_2d(win_s + b * FLAT_SLOT_BYTES + FLAT_WIN_A * 128, &p.tmX2, &win_full[b], 0, row0 + FLAT_WIN_A);
      if (has_res) tma_load_2d(win_s + b * FLAT_SLOT_BYTES + FLAT_WIN_BYTES, &p.tmR, &win_full[b], 0, tile * 128);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (converged warp, elected lane)
    constexpr uint32_t idesc = umma_idesc(CADRE_ENC_FP16 ? 0u : 1u, 0, 0, 128, 64);
    pdl_wait();
    mbar_wait(w_full, 0);
    const uint32_t w_addr = smem_u32(w_s);
    int lt = 0;
    long long d_te = 0, d_wf = 0, d_is = 0;
    const long long tstart = p.dbg ? clock64() : 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++lt) {
      const int b = lt & 1;
      const uint32_t ph = (lt >> 1) & 1;
      const int wb = lt % FLAT_NWIN;
      const long long t0 = p.dbg ? clock64() : 0;
      mbar_wait(&tempty[b], ph ^ 1);
      const long long t1 = p.dbg ? clock64() : 0;
      mbar_wait(&win_full[wb], (lt / FLAT_NWIN) & 1);
      const long long t2 = p.dbg ? clock64() : 0;
      tc_fence_after();
      const uint32_t win_addr = smem_u32(win_s + wb * FLAT_SLOT_BYTES);
      const uint32_t tacc = tmem_base + b * 64;
      if (elect_one()) {
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
      if (has_res) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint64_t da = umma_smem_desc(win_addr + FLAT_WIN_BYTES + k * 32, 16, 1024, 2);
          const uint64_t db = umma_smem_desc(w_addr + 9 * 8192 + k * 32, 16, 1024, 2);
          tc_mma_f16(tacc, da, db, idesc, 1);
        }
      }
      tc_commit(&win_empty[wb]);
      tc_commit(&tfull[b]);
      }
      __syncwarp();
      if (p.dbg) d_te += t1 - t0, d_wf += t2 - t1, d_is += clock64() - t2;
    }
    if (p.dbg && lane == 0) {
      long long* d = p.dbg + blockIdx.x * 16;
      d[2] += d_te, d[3] += d_wf, d[4] += d_is, d[5] += clock64() - tstart, d[10] += lt;
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------ epilogue: 8 warps, 2 per TMEM lane quarter
    // (with 4 warps the ~600 instructions per thread and tile were the bottleneck: 3.3 k cycles per tile
    //  against 1.7 k cycles of MMA; measured on B200)
    pdl_wait();
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
      const long long e0 = (p.dbg && leader) ? clock64() : 0;
      if (leader) tma_store_wait_read1();   // the store of tile lt-2 (same staging buffer) has been read out
      epi_bar_sync256();
      const long long e1 = (p.dbg && leader) ? clock64() : 0;
      mbar_wait(&tfull[b], ph);
      const long long e2 = (p.dbg && leader) ? clock64() : 0;
      tc_fence_after();
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + b * 64 + half * 32 + (static_cast<uint32_t>(q * 32) << 16), r);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[b]);
      uint8_t* rowp = out_s + b * (128 * 128) + row * 128;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[8 * j + i]) + bias_r[8 * j + i];
        uint4 u;
        if (p.act == ACT_RELU) {
          u.x = enc_pack2_relu(v[0], v[1]), u.y = enc_pack2_relu(v[2], v[3]);
          u.z = enc_pack2_relu(v[4], v[5]), u.w = enc_pack2_relu(v[6], v[7]);
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
        tma_store_2d(&p.tmY, out_s + b * (128 * 128), 0, tile * 128);
        tma_store_commit();
        if (p.dbg) {
          long long* d = p.dbg + blockIdx.x * 16;
          d[6] += e1 - e0, d[7] += e2 - e1, d[8] += clock64() - e2;
        }
      }
    }
    if (leader) tma_store_wait_all();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 128);
}

}  // namespace cadre

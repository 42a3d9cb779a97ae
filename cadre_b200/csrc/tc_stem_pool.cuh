// Fused ResNet stem: conv7x7/s2/p3 (Cin = 4, bias + folded BN) -> ReLU -> MaxPool2d(3, s2, p1)
// (danet_blocks/resnet.py:169-172), writing the pooled map straight into the zero-bordered layer1 input
// [B][38][66][64]. The 72x128x64 stem activation (1.18 MB / frame) never reaches HBM.
//
// Implicit GEMM as in tc_gemm.cuh MODE_STEM: conv row oh = sum_{j<4} A[rp = oh + j] * W[j], where
// A[rp] (128 pixels x 128 B, one 4-D TMA box over the row-pair interleaved image) depends only on the input
// row pair rp. A persistent CTA therefore walks the conv rows of a unit (image, block of pooled rows) in order and
// keeps the A tiles in a shared-memory ring: each new conv row needs ONE new 16 KB tile instead of four
// (L2 traffic / 3.5), the four 8 KB weight tiles stay resident. Eight epilogue warps drain the double-buffered
// TMEM accumulator (bias, ReLU, fp16) into a 3-row ring and emit one pooled row for every second conv row.
#pragma once
#include "tc_persist.cuh"

namespace cadre {

struct StemPoolParams {
  CUtensorMap tmX;   // row-pair interleaved image: dims {64, 128, 75, B}, box {64, 128, 1, 1}
  CUtensorMap tmW;   // [64][256] weights, box {64, 64}
  const float* bias; // [64]
  enc_t* out;        // [B][38][66][64] zero-bordered
  int B, pool_rows;  // pooled rows per unit (divides 36)
  int num_units;
  long long* dbg;    // optional per-CTA cycle counters [16] (cadre_debug_clk), nullptr in production
};

constexpr int SP_RING = 8;                                 // A tiles in flight (>= 4 live + prefetch)
constexpr int SP_A_BYTES = 128 * 128;
constexpr int SP_W_BYTES = 4 * 64 * 128;
constexpr int SP_ROW_BYTES = 128 * 128;                    // one conv row: 128 px x 64 ch fp16
constexpr int SP_SLOTS = 8;                                // 64-column accumulator slots (conv rows in flight): all of TMEM
constexpr int SP_SMEM = SP_RING * SP_A_BYTES + SP_W_BYTES + 3 * SP_ROW_BYTES + 48 * 8 + 16 + 1024;

__device__ __forceinline__ void sp_epi_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

__global__ void __launch_bounds__(320, 1) tc_stem_pool_kernel(const __grid_constant__ StemPoolParams p) {
  pdl_trigger();
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_s = smem;                                   // SP_RING x 16 KB
  uint8_t* w_s = a_s + SP_RING * SP_A_BYTES;             // 4 x 8 KB
  uint8_t* row_s = w_s + SP_W_BYTES;                     // 3 x 16 KB conv-row ring (128B-swizzled pixels)
  uint64_t* a_full = reinterpret_cast<uint64_t*>(row_s + 3 * SP_ROW_BYTES);
  uint64_t* a_empty = a_full + SP_RING;
  uint64_t* w_full = a_empty + SP_RING;
  uint64_t* tfull = w_full + 1;          // [SP_SLOTS]
  uint64_t* tempty = tfull + SP_SLOTS;   // [SP_SLOTS]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + SP_SLOTS);
  __shared__ float s_bias[64];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < 64) s_bias[threadIdx.x] = p.bias[threadIdx.x];
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmX);
    tma_prefetch_desc(&p.tmW);
    for (int i = 0; i < SP_RING; ++i) {
      mbar_init(&a_full[i], 1);
      mbar_init(&a_empty[i], 1);
    }
    mbar_init(w_full, 1);
    for (int i = 0; i < SP_SLOTS; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 8);  // one arrival per epilogue warp
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 64 * SP_SLOTS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int units_per_img = 36 / p.pool_rows;

  // unit -> image, first / last conv row (the row above the first pooled row is recomputed as a halo)
  auto unit_rows = [&](int unit, int& img, int& r0, int& c0, int& c1) {
    img = unit / units_per_img;
    r0 = (unit - img * units_per_img) * p.pool_rows;
    c0 = r0 == 0 ? 0 : 2 * r0 - 1;
    c1 = 2 * (r0 + p.pool_rows - 1) + 1;
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (converged warp, elected lane)
    if (elect_one()) {
      mbar_expect_tx(w_full, SP_W_BYTES);
      for (int j = 0; j < 4; ++j) tma_load_2d(w_s + (3 - j) * 8192, &p.tmW, w_full, j * 64, 0);  // W_3 first
    }
    __syncwarp();
    pdl_wait();   // weights / bias are never written by a stream predecessor; the packed input is
    int seq = 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      int img, r0, c0, c1;
      unit_rows(unit, img, r0, c0, c1);
      for (int rp = c0; rp <= c1 + 3; ++rp, ++seq) {
        const int slot = seq % SP_RING;
        const uint32_t ph = (seq / SP_RING) & 1;
        const long long t0 = p.dbg ? clock64() : 0;
        mbar_wait(&a_empty[slot], ph ^ 1);
        if (p.dbg && lane == 0) p.dbg[blockIdx.x * 16 + 1] += clock64() - t0;
        if (elect_one()) {
          mbar_expect_tx(&a_full[slot], SP_A_BYTES);
          tma_load_4d(a_s + slot * SP_A_BYTES, &p.tmX, &a_full[slot], 0, 0, rp, img);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (converged warp, elected lane)
    // SCATTER form: conv row oh = sum_j A[oh + j] W_j, so A tile rp contributes W_j-wise to the four rows
    // rp-3 .. rp. With the weights stacked as one [W_3 | W_2 | W_1 | W_0] operand (N = 256) and the rows'
    // accumulators in CONSECUTIVE 64-column TMEM slots (slot = row mod 8), ONE tcgen05.mma per k-step serves all
    // four rows: the A tile is read from shared memory once instead of four times, and a 128x256x16 MMA keeps the
    // tensor pipe busy for 128 cycles where four N = 64 MMAs were operand-read bound (~60 cycles each, measured).
    // The epilogue hands every slot back ZEROED, so all MMAs accumulate; a slot range that wraps around the ring
    // is issued as two MMAs, and the first / last tiles of a unit use the sub-range of W blocks whose rows exist.
    mbar_wait(w_full, 0);
    const uint32_t w_addr = smem_u32(w_s), a_addr0 = smem_u32(a_s);
    int seq = 0;    // A tiles consumed (ring position)
    int lt0 = 0;    // global index of the unit's first conv row (accumulator ring position)
    long long d_te = 0, d_wf = 0, d_is = 0;
    const long long tstart = p.dbg ? clock64() : 0;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      int img, r0, c0, c1;
      unit_rows(unit, img, r0, c0, c1);
      const int nrows = c1 - c0 + 1;
      for (int ti = 0; ti < nrows + 3; ++ti, ++seq) {   // tile rp = c0 + ti feeds local rows ti-3 .. ti
        const int jhi = ti < 3 ? ti : 3;                      // oldest row it feeds: ti - jhi
        const int jlo = ti > nrows - 1 ? ti - (nrows - 1) : 0;
        const int first = ti - jhi, cnt = jhi - jlo + 1;      // local rows first .. first+cnt-1 <- blocks 3-jhi ..
        const long long t0 = p.dbg ? clock64() : 0;
        if (jlo == 0) {   // row ti is new: its slot must have been drained and zeroed by the epilogue
          const int lt = lt0 + ti;
          mbar_wait(&tempty[lt % SP_SLOTS], (lt / SP_SLOTS) & 1);
        }
        const long long t1 = p.dbg ? clock64() : 0;
        const int aslot = seq % SP_RING;
        mbar_wait(&a_full[aslot], (seq / SP_RING) & 1);
        const long long t2 = p.dbg ? clock64() : 0;
        tc_fence_after();
        if (elect_one()) {
          const int s0 = (lt0 + first) % SP_SLOTS;
          const int n1 = cnt < SP_SLOTS - s0 ? cnt : SP_SLOTS - s0;   // rows before the ring wraps
          const uint32_t b0 = w_addr + (3 - jhi) * 8192;
          const uint32_t idesc1 = umma_idesc(CADRE_ENC_FP16 ? 0u : 1u, 0, 0, 128, 64 * n1);
          const uint32_t idesc2 = umma_idesc(CADRE_ENC_FP16 ? 0u : 1u, 0, 0, 128, 64 * (cnt - n1));
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t da = umma_smem_desc(a_addr0 + aslot * SP_A_BYTES + k * 32, 16, 1024, 2);
            tc_mma_f16(tmem_base + s0 * 64, da, umma_smem_desc(b0 + k * 32, 16, 1024, 2), idesc1, 1);
            if (cnt > n1)
              tc_mma_f16(tmem_base, da, umma_smem_desc(b0 + n1 * 8192 + k * 32, 16, 1024, 2), idesc2, 1);
          }
          tc_commit(&a_empty[aslot]);
          if (ti >= 3) tc_commit(&tfull[(lt0 + ti - 3) % SP_SLOTS]);   // row ti-3 has all four contributions
        }
        __syncwarp();
        if (p.dbg) d_te += t1 - t0, d_wf += t2 - t1, d_is += clock64() - t2;
      }
      lt0 += nrows;
    }
    if (p.dbg && lane == 0) {
      long long* d = p.dbg + blockIdx.x * 16;
      d[2] += d_te, d[3] += d_wf, d[4] += d_is, d[5] += clock64() - tstart, d[10] += lt0;
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------ epilogue: 8 warps, 2 per TMEM lane quarter
    pdl_wait();
    const int ew = warp - 2;            // 0..7
    const int q = warp & 3;             // TMEM lane quarter
    const int half = ew >> 2;           // which 32 of the 64 channels
    const int px = q * 32 + lane;       // conv pixel (0..127)
    const int et = threadIdx.x - 64;    // 0..255
    float bias_r[32];                   // this thread always handles the same 32 channels
#pragma unroll
    for (int i = 0; i < 32; ++i) bias_r[i] = s_bias[half * 32 + i];
    // all accumulator slots start zeroed (TMEM is not cleared by the allocator); this is tempty's phase 0
    for (int sl = 0; sl < SP_SLOTS; ++sl)
      tmem_st_zero_32x32(tmem_base + sl * 64 + half * 32 + (static_cast<uint32_t>(q * 32) << 16));
    tmem_st_wait();
    tc_fence_before();
    __syncwarp();
    if (lane == 0)
      for (int sl = 0; sl < SP_SLOTS; ++sl) mbar_arrive(&tempty[sl]);
    int lt = 0, npool = 0;
    uint32_t prev[16], vm[16];          // packed pairs: the odd conv row above, the running vertical maximum
#pragma unroll
    for (int i = 0; i < 16; ++i) prev[i] = vm[i] = 0u;
    for (int unit = blockIdx.x; unit < p.num_units; unit += gridDim.x) {
      int img, r0, c0, c1;
      unit_rows(unit, img, r0, c0, c1);
      bool have_prev = false;
      for (int oh = c0; oh <= c1; ++oh, ++lt) {
        const int as = lt % SP_SLOTS;
        const uint32_t aph = (lt / SP_SLOTS) & 1;
        const bool dbgl = p.dbg && et == 0;
        const long long e0 = dbgl ? clock64() : 0;
        mbar_wait(&tfull[as], aph);
        const long long e1 = dbgl ? clock64() : 0;
        tc_fence_after();
        uint32_t r[32];
        const uint32_t tslot = tmem_base + as * 64 + half * 32 + (static_cast<uint32_t>(q * 32) << 16);
        tmem_ld_32x32(tslot, r);
        tmem_ld_wait();
        tmem_st_zero_32x32(tslot);   // hand the slot back zeroed: every MMA accumulates
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[as]);
        // bias + ReLU + fp16 pack in registers. The 3x3/s2 max-pool is separable and this thread sees the SAME pixel
        // column in every conv row, so the vertical max (rows 2r-1, 2r, 2r+1) stays in registers; only the vertically
        // reduced row goes to shared memory (once per pooled row) for the horizontal 3-max + stride-2 pick:
        // 41 KB of shared-memory traffic per pooled row instead of 106 KB (the kernel was bound by that traffic).
#if CADRE_ENC_FP16
        typedef __half2 enc2_t;
#else
        typedef __nv_bfloat162 enc2_t;
#endif
        uint32_t cur[16];
#pragma unroll
        for (int i = 0; i < 16; ++i)
          cur[i] = enc_pack2_relu(__uint_as_float(r[2 * i]) + bias_r[2 * i], __uint_as_float(r[2 * i + 1]) + bias_r[2 * i + 1]);
        const long long e2 = dbgl ? clock64() : 0;
        if (dbgl) p.dbg[blockIdx.x * 16 + 7] += e1 - e0, p.dbg[blockIdx.x * 16 + 8] += e2 - e1;
        bool emit = false;
        if ((oh & 1) == 0) {   // row 2r: start the pooled row from the kept odd row above it (absent for r == 0)
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            enc2_t m = *reinterpret_cast<enc2_t*>(&cur[i]);
            if (have_prev) m = __hmax2(m, *reinterpret_cast<enc2_t*>(&prev[i]));
            vm[i] = *reinterpret_cast<uint32_t*>(&m);
          }
        } else if (((oh - 1) >> 1) >= r0) {   // row 2r+1 completes pooled row r
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            enc2_t m = __hmax2(*reinterpret_cast<enc2_t*>(&vm[i]), *reinterpret_cast<enc2_t*>(&cur[i]));
            vm[i] = *reinterpret_cast<uint32_t*>(&m);
          }
          emit = true;
        }
        if (oh & 1) {   // an odd row is also row 2(r+1)-1 of the next pooled row (the unit's halo row only that)
#pragma unroll
          for (int i = 0; i < 16; ++i) prev[i] = cur[i];
          have_prev = true;
        }
        if (emit) {
          const int r = (oh - 1) >> 1;
          uint8_t* vrow = row_s + (npool & 1) * SP_ROW_BYTES;   // double buffered: one barrier per pooled row
          ++npool;
          uint8_t* rowp = vrow + px * 128;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int chunk = half * 4 + j;
            *reinterpret_cast<uint4*>(rowp + ((chunk ^ (px & 7)) << 4)) =
                make_uint4(vm[4 * j], vm[4 * j + 1], vm[4 * j + 2], vm[4 * j + 3]);
          }
          sp_epi_bar();
          enc_t* orow = p.out + ((static_cast<long long>(img) * 38 + r + 1) * 66 + 1) * 64;
#pragma unroll
          for (int it = 0; it < 2; ++it) {
            const int item = et + it * 256;      // 0..511: pooled pixel pw, 16-byte channel chunk
            const int pw = item >> 3, chunk = item & 7;
            uint4 u[3];
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {   // column -1 is padding: re-reading column 0 leaves the max unchanged
              const int x = (2 * pw + dx >= 0) ? 2 * pw + dx : 0;
              u[dx + 1] = *reinterpret_cast<const uint4*>(vrow + x * 128 + ((chunk ^ (x & 7)) << 4));
            }
            uint4 o;
            enc2_t* m2 = reinterpret_cast<enc2_t*>(&o);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              m2[i] = __hmax2(__hmax2(reinterpret_cast<const enc2_t*>(&u[0])[i], reinterpret_cast<const enc2_t*>(&u[1])[i]),
                              reinterpret_cast<const enc2_t*>(&u[2])[i]);
            *reinterpret_cast<uint4*>(orow + pw * 64 + chunk * 8) = o;
          }
          if (dbgl) p.dbg[blockIdx.x * 16 + 9] += clock64() - e2;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 64 * SP_SLOTS);
}

}  // namespace cadre

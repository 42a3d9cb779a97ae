// Persistent LSTM recurrence kernels of the PPO update (ppo_agent/models.py:139-152: eight sequential nn.LSTMCell
// steps per row; ppo_agent/agent.py:170-182 evaluates them for every command expert).
//
// The recurrent part of a step, gates += h_{t-1} W_hh^T, used to be one grid launch per step (8 forward, 7 + 8
// backward): every launch re-read its 128 x 530 fp32 slice of W_hh from L2 and paid launch + pipeline fill. Here ONE
// kernel per direction keeps the recurrent weights resident in shared memory for all eight steps:
//   * grid = 8 experts x 17 slices = 136 CTAs (<= 148 SMs, one CTA per SM, all co-resident);
//   * forward  : CTA (e, j) owns gate columns [128 j, 128 j + 128) = hidden units [32 j, 32 j + 32): its W_hh rows as
//                fp16 [128 x 576] (147 KB, K zero-padded 530 -> 576) in the canonical 128B-swizzled K-major layout;
//   * backward : CTA (e, j) owns hidden units [32 j, 32 j + 32) of dh: W_hh[:, units] transposed, fp16 [32 x 2176];
//     (both slices arrive by TMA from fp16 copies of W_hh that wcvt_kernel (ppo.cu) refreshes at the start of every
//     update, concurrently with the routing kernels);
//   * per step the A operand (h_{t-1}, resp. dG_t, all rows of the expert, fp16) streams from L2 through a TMA ring,
//     tcgen05.mma (kind::f16, fp32 accumulation in TMEM) produces the CTA's slice, the epilogue warps apply the LSTM
//     cell (resp. its derivative) and write the next step's operand slice back to global memory;
//   * the 17 CTAs of an expert exchange their slices through L2: release-increment of a per-expert counter after the
//     slice is written, acquire-poll before the next step's first TMA load (generic -> async proxy fences on both
//     sides). Different experts never wait for each other.
// fp16 operands carry the same 11-bit significand as the TF32 path they replace; accumulation, cell state, gate
// activations and every tensor kept for the weight-gradient GEMMs stay fp32. Backward operands are scaled by a power
// of two (params.scale) so that small gradients stay in fp16's normal range; conversions saturate.
#pragma once
#include <cuda_fp16.h>

#include "internal.h"
#include "ppo_layout.h"
#include "ptx.cuh"

namespace cadre {

constexpr int LS_SLICES = 17;                 // ceil(2120 / 128) = ceil(530 / 32)
constexpr int LS_A_STAGE = 128 * 128;         // one A k-block: 128 rows x 64 halves
constexpr int LS_LDH16 = 544;                 // row pitch (halves) of the fp16 x / h tensors  [E][cap][9][544]
constexpr int LS_LDG16 = 2176;                // row pitch (halves) of the fp16 dG tensor      [E][cap][9][2176]
constexpr int LS_STG_LD = 33;                 // fp32 staging pitch (conflict-free row and column access)
constexpr unsigned LS_SPIN_LIMIT = 1u << 22;  // bounded polling: a lost hand-off raises an error flag instead of hanging

// forward
constexpr int LSF_KB = 9;                     // 64-wide k-blocks over K = 530 (padded to 576)
constexpr int LSF_STAGES = 3;
constexpr int LSF_W_BYTES = LSF_KB * 128 * 128;
constexpr int LSF_STG_BYTES = 8 * 32 * LS_STG_LD * 4;
constexpr int LSF_SMEM = LSF_W_BYTES + LSF_STAGES * LS_A_STAGE + LSF_STG_BYTES + 256 + 1024;
constexpr int LSF_THREADS = 64 + 256;
// backward
constexpr int LSB_KB = 34;                    // 64-wide k-blocks over K = 2120 (padded to 2176)
constexpr int LSB_STAGES = 4;
constexpr int LSB_W_BYTES = LSB_KB * 32 * 128;
constexpr int LSB_STG_BYTES = 4 * 32 * LS_STG_LD * 4;
constexpr int LSB_SMEM = LSB_W_BYTES + LSB_STAGES * LS_A_STAGE + LSB_STG_BYTES + 256 + 1024;
constexpr int LSB_THREADS = 64 + 256;
constexpr int LSB_NBUF = 4;                   // TMEM accumulator buffers (32 columns each)

struct LstmFwdParams {
  CUtensorMap tmH;        // H16 as {k = 530, row = cap, slot = 9, expert = 8}, box {64, 128, 1, 1}, 128B swizzle
  CUtensorMap tmW;        // W_hh as fp16 [E][G][544] {k = 530, gate row = 2120, expert}, box {64, 128, 1}
  const float* XP9;       // [E][cap][9][G]   x_t W_ih^T + b_ih + b_hh
  __half* G16;            // [E][cap][9][G]   gate activations (i, f, g, o per unit), kept for the backward pass
  float* C9;              // [E][cap][9][LDF] slot t holds c_{t-1}
  float* H8;              // [E][cap][LDF]    h_8 in fp32: input of the actor / critic heads
  __half* H16;            // [E][cap][9][LS_LDH16] slot t holds h_{t-1} (slot 0 written by the gather): the recurrent
                          // GEMM's A operand of step t and the W_hh weight-gradient GEMM's B operand
  const int* counts;      // [E] routed rows per expert
  unsigned* sync;         // [E] arrival counters (zero at launch), [E] = error flag
  int cap;
  long long* dbg;         // optional clock64 stamps [CTA][9][8] (cadre_debug_clk), nullptr in production
};

struct LstmBwdParams {
  CUtensorMap tmDG;       // dG16 as {k = 2176, row = cap, slot = 9, expert = 8}, box {64, 128, 1, 1}
  CUtensorMap tmWT;       // W_hh^T as fp16 [E][544][2176] {k = gate row 2120, unit = 530, expert}, box {64, 32, 1}
  const __half* G16;
  const float* C9;
  __half* dG16;           // [E][cap][9][LS_LDG16] slot t holds scale * dG_t (d loss / d gate pre-activations): A operand
                          // of the recurrent dgrad GEMM of step t and of both LSTM weight-gradient GEMMs
  const float* dH8;       // [E][cap][LDF]   d loss / d h_8 (from the first-layer dgrad GEMM)
  float* dC;              // [E][cap][LDF]   running d loss / d c
  const int* counts;
  unsigned* sync;         // [E] counters (zero at launch), [E] = error flag
  int cap;
  float scale, inv_scale;
  float* grads;           // flat gradient buffer: the LSTM bias gradients (column sums of dG over rows and steps) are
                          // accumulated by the epilogue and written to the expert's b_ih / b_hh
  long long* dbg;
};

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }

// wait until *ctr >= target (one polling lane, the warp re-converges afterwards); sets *err on timeout
__device__ __forceinline__ void wait_counter(const unsigned* ctr, unsigned target, unsigned* err, int lane) {
  if (lane == 0) {
    unsigned spins = 0;
    while (ld_acquire_u32(ctr) < target) {
      __nanosleep(32);
      if (++spins > LS_SPIN_LIMIT) {
        atomicExch(err, 1u);
        break;
      }
    }
  }
  __syncwarp();
  fence_proxy_async_all();   // the slices acquired above were written through the generic proxy; TMA reads them next
}

__device__ __forceinline__ float ls_sigmoid(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float ls_tanh(float x) { return 1.0f - __fdividef(2.0f, __expf(2.0f * x) + 1.0f); }
__device__ __forceinline__ unsigned short f2h_sat_bits(float x) {
  unsigned short r;
  asm("cvt.rn.satfinite.f16.f32 %0, %1;" : "=h"(r) : "f"(x));
  return r;
}

// =========================================================================================================== forward
__global__ void __launch_bounds__(LSF_THREADS, 1) lstm_seq_fwd_kernel(const __grid_constant__ LstmFwdParams p) {
  using namespace ppo;
  extern __shared__ uint8_t ls_raw[];
  uint8_t* smem = ls_raw + ((1024u - (smem_u32(ls_raw) & 1023u)) & 1023u);
  uint8_t* w_s = smem;                                   // LSF_KB tiles [128 n x 64 k]
  uint8_t* a_s = w_s + LSF_W_BYTES;                      // LSF_STAGES tiles [128 rows x 64 k]
  float* stg_all = reinterpret_cast<float*>(a_s + LSF_STAGES * LS_A_STAGE);
  uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stg_all) + LSF_STG_BYTES);
  uint64_t* empty = full + LSF_STAGES;
  uint64_t* acc_full = empty + LSF_STAGES;               // [2]
  uint64_t* acc_empty = acc_full + 2;                    // [2]
  uint64_t* w_full = acc_empty + 2;                      // resident weights have landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.y, j = blockIdx.x;
  const int n0 = j * 128;                                // first gate column (= W_hh row) of this CTA
  long long* dbg = p.dbg ? p.dbg + static_cast<long long>(e * LS_SLICES + j) * 72 : nullptr;   // [9][8] stamps
  pdl_trigger();
  if (dbg && threadIdx.x == 0) dbg[64] = clock64();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmH);
    tma_prefetch_desc(&p.tmW);
    for (int s = 0; s < LSF_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 8);      // one arrival per epilogue warp
    }
    mbar_init(w_full, 1);
    fence_mbar_init();
    // ---- resident weights: rows [n0, n0 + 128) of the fp16 copy of W_hh[e] (written by wcvt_kernel at the start of
    // this update; the host joins that kernel's stream before any launch of the forward pass, so it is complete even
    // though this runs before griddepcontrol.wait). Nine TMA boxes land directly in the swizzled K-major layout;
    // k >= 530 and gate rows >= 2120 are filled with zeros by the TMA unit.
    mbar_expect_tx(w_full, LSF_W_BYTES);
    for (int kb = 0; kb < LSF_KB; ++kb) tma_load_3d(w_s + kb * (128 * 128), &p.tmW, w_full, kb * 64, n0, e);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  pdl_wait();                           // counts, XP9, H16 buffer 0, the zeroed counters: all from earlier launches
  if (dbg && threadIdx.x == 0) dbg[66] = clock64();
  const int count = p.counts[e];
  const int n_mt = (count + 127) >> 7;
  unsigned* ctr = p.sync + e;
  unsigned* err = p.sync + E;

  if (count > 0) {
    if (warp == 0) {
      // ------------------------------------------------------------------ TMA producer: h_{t-1}, all rows of expert e
      uint32_t it = 0;
      for (int t = 0; t < 8; ++t) {
        if (dbg && lane == 0) dbg[t * 8 + 0] = clock64();
        if (t > 0) wait_counter(ctr, static_cast<unsigned>(LS_SLICES * t), err, lane);
        if (dbg && lane == 0) dbg[t * 8 + 1] = clock64();
        for (int mt = 0; mt < n_mt; ++mt)
          for (int kb = 0; kb < LSF_KB; ++kb, ++it) {
            const int s = it % LSF_STAGES;
            mbar_wait(&empty[s], ((it / LSF_STAGES) & 1) ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&full[s], LS_A_STAGE);
              tma_load_4d(a_s + s * LS_A_STAGE, &p.tmH, &full[s], kb * 64, mt * 128, t, e);
            }
            __syncwarp();
          }
        if (dbg && lane == 0) dbg[t * 8 + 2] = clock64();
      }
    } else if (warp == 1) {
      // ------------------------------------------------------------------ MMA issuer
      constexpr uint32_t idesc = umma_idesc(0u, 0, 0, 128, 128);   // fp16 x fp16 -> fp32, both K-major
      uint32_t it = 0, tile = 0;
      mbar_wait(w_full, 0);
      if (dbg && lane == 0) dbg[65] = clock64();             // weights resident
      for (int t = 0; t < 8; ++t)
        for (int mt = 0; mt < n_mt; ++mt, ++tile) {
          const uint32_t buf = tile & 1;
          mbar_wait(&acc_empty[buf], ((tile >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * 128;
          for (int kb = 0; kb < LSF_KB; ++kb, ++it) {
            const int s = it % LSF_STAGES;
            mbar_wait(&full[s], (it / LSF_STAGES) & 1);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(a_s + s * LS_A_STAGE);
            const uint32_t b_addr = smem_u32(w_s + kb * (128 * 128));
            if (elect_one()) {
              const int nk = kb == LSF_KB - 1 ? 2 : 4;   // k >= 544 is zero on both sides
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (k < nk)
                  tc_mma_f16(d_tmem, umma_smem_desc(a_addr + k * 32, 16, 1024, 2),
                             umma_smem_desc(b_addr + k * 32, 16, 1024, 2), idesc, (kb | k) != 0);
              tc_commit(&empty[s]);
              if (kb == LSF_KB - 1) tc_commit(&acc_full[buf]);
            }
            __syncwarp();
          }
          if (dbg && lane == 0) dbg[t * 8 + 3] = clock64();
        }
    } else {
      // ------------------------------------------------------------------ epilogue: LSTM cell (8 warps)
      // Order of the global stores: the fp16 h slice the OTHER CTAs are waiting for goes out first and is published;
      // gates / cell state (read by later kernels and by this thread only) follow, so their store traffic overlaps
      // the next step's hand-off. The epilogue is instruction-issue bound (4096 cells per CTA and step), so every
      // address is a per-thread base pointer plus a compile-time multiple of the 4-row stride, and row validity is
      // two bit masks.
      const int q = warp & 3, hf = (warp - 2) >> 2;      // TMEM lane quarter, column half
      float* stg = stg_all + (warp - 2) * 32 * LS_STG_LD;
      const int rsub = lane >> 3, ul = lane & 7;         // coalesced pass: 4 rows x 8 units per instruction
      const int colA = n0 + hf * 64 + 4 * ul;            // gate column of this lane's unit in pass 0 (+32 in pass 1)
      const bool colok[2] = {colA < G, colA + 32 < G};
      constexpr long long S_G = 4LL * 9 * G, S_C = 4LL * 9 * LDF, S_H16 = 4LL * 9 * LS_LDH16, S_H8 = 4LL * LDF;
      const float* stg_rd = stg + rsub * LS_STG_LD + 4 * ul;
      uint32_t tile = 0;
      for (int t = 0; t < 8; ++t) {
        for (int mt = 0; mt < n_mt; ++mt, ++tile) {
          const uint32_t buf = tile & 1;
          const int m0 = mt * 128 + q * 32 + rsub;       // this lane's first row; row i is m0 + 4 i
          const long long rg0 = static_cast<long long>(e) * p.cap + m0;
          // bit i: row m0 + 4 i is a routed row / has to be written (zeros up to the next multiple of 32 rows)
          uint32_t ok_mask = 0, wr_mask = 0;
          {
            const int wr_end = (count + 31) & ~31;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              ok_mask |= (m0 + 4 * i < count ? 1u : 0u) << i;
              wr_mask |= (m0 + 4 * i < wr_end ? 1u : 0u) << i;
            }
          }
          const float* xp_p = p.XP9 + (rg0 * 9 + t) * G + colA;
          float* c_p = p.C9 + (rg0 * 9 + t) * LDF + (colA >> 2);
          // x-part pre-activations and c_{t-1} of both column passes: issued BEFORE waiting for the accumulator, so
          // that their L2 / HBM latency hides behind the hand-off wait and the MMAs of this step
          float4 x4[2][8];
          float cp[2][8];       // becomes c_t in place
          uint2 gb[2][8];       // gate activations (i, f, g, o) as fp16
          float h8[2][8];
#pragma unroll
          for (int pass = 0; pass < 2; ++pass)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const bool ok = ((ok_mask >> i) & 1u) && colok[pass];
              x4[pass][i] = ok ? *reinterpret_cast<const float4*>(xp_p + i * S_G + pass * 32) : make_float4(0.f, 0.f, 0.f, 0.f);
              cp[pass][i] = ok ? c_p[i * S_C + pass * 8] : 0.f;
            }
          mbar_wait(&acc_full[buf], (tile >> 1) & 1);
          tc_fence_after();
          if (dbg && threadIdx.x == 64) dbg[t * 8 + 4] = clock64();
          unsigned short* h16_p = reinterpret_cast<unsigned short*>(p.H16) + (rg0 * 9 + t + 1) * LS_LDH16 + (colA >> 2);
#pragma unroll
          for (int pass = 0; pass < 2; ++pass) {
            uint32_t r[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * 128 + hf * 64 + pass * 32, r);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 32; ++c) stg[lane * LS_STG_LD + c] = __uint_as_float(r[c]);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              float gi = 0.f, gf = 0.f, gg = 0.f, go = 0.f, cn = 0.f, h = 0.f;
              if (((ok_mask >> i) & 1u) && colok[pass]) {
                const float* a = stg_rd + i * 4 * LS_STG_LD;
                gi = ls_sigmoid(a[0] + x4[pass][i].x);
                gf = ls_sigmoid(a[1] + x4[pass][i].y);
                gg = ls_tanh(a[2] + x4[pass][i].z);
                go = ls_sigmoid(a[3] + x4[pass][i].w);
                cn = fmaf(gf, cp[pass][i], gi * gg);
                h = go * ls_tanh(cn);
              }
              // gates in (-1, 1): fp16 keeps an absolute error <= 2.5e-4
              gb[pass][i].x = static_cast<uint32_t>(f2h_sat_bits(gi)) | (static_cast<uint32_t>(f2h_sat_bits(gf)) << 16);
              gb[pass][i].y = static_cast<uint32_t>(f2h_sat_bits(gg)) | (static_cast<uint32_t>(f2h_sat_bits(go)) << 16);
              cp[pass][i] = cn;
              h8[pass][i] = h;
              if (((wr_mask >> i) & 1u) && colok[pass]) h16_p[i * S_H16 + pass * 8] = f2h_sat_bits(h);
            }
            __syncwarp();
          }
          tc_fence_before();
          if (lane == 0) mbar_arrive(&acc_empty[buf]);
          if (mt == n_mt - 1) {
            // publish h_t of this slice (all row tiles): barrier of the epilogue warps, then ONE thread releases the
            // counter. red.release.gpu orders every write this thread observed through the barrier (cumulativity)
            // before the increment; the generic -> async proxy fence covers the consumers' TMA reads.
            if (dbg && threadIdx.x == 64) dbg[t * 8 + 5] = clock64();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (threadIdx.x == 64) {
              fence_proxy_async_all();
              red_release_add_u32(ctr, 1u);
            }
            if (dbg && threadIdx.x == 64) dbg[t * 8 + 6] = clock64();
          }
          // the tensors kept for the backward pass (and h_8, the input of the heads)
          __half* g16_p = p.G16 + (rg0 * 9 + t) * G + colA;
          float* h8_p = p.H8 + rg0 * LDF + (colA >> 2);
#pragma unroll
          for (int pass = 0; pass < 2; ++pass)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (!(((wr_mask >> i) & 1u) && colok[pass])) continue;
              *reinterpret_cast<uint2*>(g16_p + i * S_G + pass * 32) = gb[pass][i];
              c_p[(i * S_C + pass * 8) + LDF] = cp[pass][i];          // slot t + 1
              if (t == 7) h8_p[i * S_H8 + pass * 8] = h8[pass][i];
            }
          if (dbg && threadIdx.x == 64) dbg[t * 8 + 7] = clock64();
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

// ========================================================================================================== backward
// Per step t = 7 .. 0 and expert:  dG_t = cell'(dh_t, dc_t; gates_t, c_{t-1}, c_t),  dc_{t-1} = dc_t * f_t,
// dh_{t-1} = dG_t W_hh  (models.py:146-151 backwards; the pointwise part is the derivative of nn.LSTMCell).
__global__ void __launch_bounds__(LSB_THREADS, 1) lstm_seq_bwd_kernel(const __grid_constant__ LstmBwdParams p) {
  using namespace ppo;
  extern __shared__ uint8_t ls_raw[];
  uint8_t* smem = ls_raw + ((1024u - (smem_u32(ls_raw) & 1023u)) & 1023u);
  uint8_t* w_s = smem;                                   // LSB_KB tiles [32 units x 64 gate rows]
  uint8_t* a_s = w_s + LSB_W_BYTES;
  float* stg_all = reinterpret_cast<float*>(a_s + LSB_STAGES * LS_A_STAGE);
  uint64_t* full = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(stg_all) + LSB_STG_BYTES);
  uint64_t* empty = full + LSB_STAGES;
  uint64_t* acc_full = empty + LSB_STAGES;               // [LSB_NBUF]
  uint64_t* acc_empty = acc_full + LSB_NBUF;             // [LSB_NBUF]
  uint64_t* w_full = acc_empty + LSB_NBUF;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int e = blockIdx.y, j = blockIdx.x;
  const int u0 = j * 32;                                 // first hidden unit of this CTA
  long long* dbg = p.dbg ? p.dbg + static_cast<long long>(e * LS_SLICES + j) * 72 : nullptr;
  pdl_trigger();
  if (dbg && threadIdx.x == 0) dbg[64] = clock64();

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmDG);
    tma_prefetch_desc(&p.tmWT);
    for (int s = 0; s < LSB_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int b = 0; b < LSB_NBUF; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], 8);      // one arrival per epilogue warp
    }
    mbar_init(w_full, 1);
    fence_mbar_init();
    // ---- resident weights: B[n = unit][k = gate row] = W_hh[e][k][u0 + n] from the transposed fp16 copy written by
    // wcvt_kernel (see the forward kernel): 34 TMA boxes of [32 units x 64 gate rows]
    mbar_expect_tx(w_full, LSB_W_BYTES);
    for (int kb = 0; kb < LSB_KB; ++kb) tma_load_3d(w_s + kb * (32 * 128), &p.tmWT, w_full, kb * 64, u0, e);
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 32 * LSB_NBUF);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  pdl_wait();
  if (dbg && threadIdx.x == 0) dbg[66] = clock64();
  const int count = p.counts[e];
  const int n_mt = (count + 127) >> 7;
  unsigned* ctr = p.sync + e;
  unsigned* err = p.sync + E;

  if (count > 0) {
    if (warp == 0) {
      // ------------------------------------------------------------------ TMA producer: dG_t, all rows of expert e
      uint32_t it = 0;
      for (int t = 7; t >= 1; --t) {
        if (dbg && lane == 0) dbg[t * 8 + 0] = clock64();
        wait_counter(ctr, static_cast<unsigned>(LS_SLICES * (8 - t)), err, lane);
        if (dbg && lane == 0) dbg[t * 8 + 1] = clock64();
        for (int mt = 0; mt < n_mt; ++mt)
          for (int kb = 0; kb < LSB_KB; ++kb, ++it) {
            const int s = it % LSB_STAGES;
            mbar_wait(&empty[s], ((it / LSB_STAGES) & 1) ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&full[s], LS_A_STAGE);
              tma_load_4d(a_s + s * LS_A_STAGE, &p.tmDG, &full[s], kb * 64, mt * 128, t, e);
            }
            __syncwarp();
          }
        if (dbg && lane == 0) dbg[t * 8 + 2] = clock64();
      }
    } else if (warp == 1) {
      // ------------------------------------------------------------------ MMA issuer: dh_{t-1}[:, units] = dG_t W_hh
      constexpr uint32_t idesc = umma_idesc(0u, 0, 0, 128, 32);
      uint32_t it = 0, tile = 0;
      mbar_wait(w_full, 0);
      if (dbg && lane == 0) dbg[65] = clock64();
      for (int t = 7; t >= 1; --t)
        for (int mt = 0; mt < n_mt; ++mt, ++tile) {
          const uint32_t buf = tile % LSB_NBUF;
          mbar_wait(&acc_empty[buf], ((tile / LSB_NBUF) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * 32;
          for (int kb = 0; kb < LSB_KB; ++kb, ++it) {
            const int s = it % LSB_STAGES;
            mbar_wait(&full[s], (it / LSB_STAGES) & 1);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(a_s + s * LS_A_STAGE);
            const uint32_t b_addr = smem_u32(w_s + kb * (32 * 128));
            if (elect_one()) {
              const int nk = kb == LSB_KB - 1 ? 1 : 4;   // gate rows >= 2128 are zero on both sides (2120 = 33*64 + 8)
#pragma unroll
              for (int k = 0; k < 4; ++k)
                if (k < nk)
                  tc_mma_f16(d_tmem, umma_smem_desc(a_addr + k * 32, 16, 1024, 2),
                             umma_smem_desc(b_addr + k * 32, 16, 1024, 2), idesc, (kb | k) != 0);
              tc_commit(&empty[s]);
              if (kb == LSB_KB - 1) tc_commit(&acc_full[buf]);
            }
            __syncwarp();
          }
          if (dbg && lane == 0) dbg[t * 8 + 3] = clock64();
        }
    } else {
      // ------------------------------------------------------------------ epilogue: LSTM cell backward (8 warps)
      // Two warps per TMEM lane quarter: each drains 16 accumulator columns into the quarter's staging tile, then owns
      // 16 of its 32 rows with lane = hidden unit. Everything the pointwise derivative needs besides dh (gates, c_{t-1},
      // c_t, running dc) is loaded BEFORE the accumulator wait; of the results only the fp16 dG slice the other CTAs
      // wait for is stored ahead of the fence + publish, the fp32 dG / dc stores follow.
      const int q = warp & 3, hf = (warp - 2) >> 2;
      float* stg = stg_all + q * 32 * LS_STG_LD;         // shared by the quarter's two warps
      const int unit = u0 + lane;
      const bool unit_ok = unit < F;
      const int qbar = 2 + q;                             // named barrier of the quarter (64 threads)
      uint32_t tile = 0;                                  // accumulator tiles consumed so far
      float4 bias_acc = make_float4(0.f, 0.f, 0.f, 0.f);  // column sums of dG over this warp's rows, all tiles and steps
      constexpr long long R_G = 9LL * G, R_C = 9LL * LDF, R_DG = 9LL * LS_LDG16;   // element strides between rows
      const int rt0 = q * 32 + hf * 16;                   // this warp's first row inside a tile
      const float* stg_rd = stg + hf * 16 * LS_STG_LD + lane;
      for (int t = 7; t >= 0; --t) {
        for (int mt = 0; mt < n_mt; ++mt) {
          const bool from_acc = t < 7;
          const int m0 = mt * 128 + rt0;                  // rows m0 .. m0 + 15
          const long long rg0 = static_cast<long long>(e) * p.cap + m0;
          uint32_t ok_mask = 0, wr_mask = 0;              // bit i: row m0 + i is routed / must be written (zeros)
          if (unit_ok) {
            const int wr_end = (count + 31) & ~31;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              ok_mask |= (m0 + i < count ? 1u : 0u) << i;
              wr_mask |= (m0 + i < wr_end ? 1u : 0u) << i;
            }
          }
          const __half* g_p = p.G16 + (rg0 * 9 + t) * G + 4 * unit;
          const float* c_p = p.C9 + (rg0 * 9 + t) * LDF + unit;
          float* dc_p = p.dC + rg0 * LDF + unit;
          float4 g4[16];                                  // gates (i, f, g, o)
          float cprev[16], ct[16], dcin[16], dh[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const bool ok = (ok_mask >> i) & 1u;
            uint2 gb = make_uint2(0u, 0u);
            if (ok) gb = *reinterpret_cast<const uint2*>(g_p + i * R_G);
            const float2 lo = __half22float2(*reinterpret_cast<const __half2*>(&gb.x));
            const float2 hi = __half22float2(*reinterpret_cast<const __half2*>(&gb.y));
            g4[i] = make_float4(lo.x, lo.y, hi.x, hi.y);
            cprev[i] = ok ? c_p[i * R_C] : 0.f;
            ct[i] = ok ? c_p[i * R_C + LDF] : 0.f;
            dcin[i] = (ok && from_acc) ? dc_p[i * LDF] : 0.f;
            dh[i] = (ok && !from_acc) ? p.dH8[(rg0 + i) * LDF + unit] : 0.f;
          }
          if (from_acc) {
            const uint32_t buf = tile % LSB_NBUF;
            mbar_wait(&acc_full[buf], (tile / LSB_NBUF) & 1);
            tc_fence_after();
            if (dbg && threadIdx.x == 64 && mt == 0) dbg[t * 8 + 4] = clock64();
            uint32_t r[16];
            tmem_ld_32x16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * 32 + hf * 16, r);
            tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 16; ++c) stg[lane * LS_STG_LD + hf * 16 + c] = __uint_as_float(r[c]);
            tc_fence_before();
            if (lane == 0) mbar_arrive(&acc_empty[buf]);  // registers hold the columns: the buffer may be refilled
            ++tile;
            asm volatile("bar.sync %0, 64;" ::"r"(qbar) : "memory");   // both halves of the quarter are staged
#pragma unroll
            for (int i = 0; i < 16; ++i) dh[i] = stg_rd[i * LS_STG_LD] * p.inv_scale;
          }
          __half* dg_p = p.dG16 + (rg0 * 9 + t) * LS_LDG16 + 4 * unit;
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float4 d = make_float4(0.f, 0.f, 0.f, 0.f);
            float dcp = 0.f;
            if ((ok_mask >> i) & 1u) {
              const float4 g = g4[i];            // i, f, g, o
              const float tc = ls_tanh(ct[i]);
              const float dho = dh[i] * g.w;
              const float dc = fmaf(dho, 1.f - tc * tc, dcin[i]);
              d = make_float4(dc * g.z * g.x * (1.f - g.x), dc * cprev[i] * g.y * (1.f - g.y),
                              dc * g.x * (1.f - g.z * g.z), dho * tc * (1.f - g.w));
              dcp = dc * g.y;
              bias_acc.x += d.x, bias_acc.y += d.y, bias_acc.z += d.z, bias_acc.w += d.w;
            }
            dcin[i] = dcp;
            if ((wr_mask >> i) & 1u) {
              uint2 hbits;
              hbits.x = static_cast<uint32_t>(f2h_sat_bits(d.x * p.scale)) |
                        (static_cast<uint32_t>(f2h_sat_bits(d.y * p.scale)) << 16);
              hbits.y = static_cast<uint32_t>(f2h_sat_bits(d.z * p.scale)) |
                        (static_cast<uint32_t>(f2h_sat_bits(d.w * p.scale)) << 16);
              *reinterpret_cast<uint2*>(dg_p + i * R_DG) = hbits;
            }
          }
          if (mt == n_mt - 1 && t > 0) {   // publish dG_t of this slice (all row tiles); no CTA waits for dG_0
            if (dbg && threadIdx.x == 64) dbg[t * 8 + 5] = clock64();
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (threadIdx.x == 64) {       // release: see the forward kernel
              fence_proxy_async_all();
              red_release_add_u32(ctr, 1u);
            }
            if (dbg && threadIdx.x == 64) dbg[t * 8 + 6] = clock64();
          } else if (from_acc) {
            asm volatile("bar.sync %0, 64;" ::"r"(qbar) : "memory");   // staging is rewritten by the next tile
          }
          if (t > 0) {   // running d loss / d c for the next (earlier) step
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if ((ok_mask >> i) & 1u) dc_p[i * LDF] = dcin[i];
          }
          if (dbg && threadIdx.x == 64) dbg[t * 8 + 7] = clock64();
        }
      }
      // LSTM bias gradients: d loss / d b_ih = d loss / d b_hh = sum over rows and steps of dG (models.py:133-137).
      // Fixed summation order (rows of a warp in sequence, then the eight warps in order): deterministic.
      float4* red = reinterpret_cast<float4*>(stg_all);   // staging is free now: [8 warps][32 lanes]
      asm volatile("bar.sync 1, 256;" ::: "memory");
      red[(warp - 2) * 32 + lane] = bias_acc;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (warp == 2 && unit_ok) {
        float4 s4 = red[lane];
#pragma unroll
        for (int w = 1; w < 8; ++w) {
          const float4 o = red[w * 32 + lane];
          s4.x += o.x, s4.y += o.y, s4.z += o.z, s4.w += o.w;
        }
        *reinterpret_cast<float4*>(p.grads + OFF_LSTM + e * LSTM_BLK + LSTM_BIH + 4 * unit) = s4;
        *reinterpret_cast<float4*>(p.grads + OFF_LSTM + e * LSTM_BLK + LSTM_BHH + 4 * unit) = s4;
      }
    }
  } else if (warp == 2) {
    // an expert without rows: its LSTM bias gradients are exact zeros
    const int unit = u0 + lane;
    if (unit < F) {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
      *reinterpret_cast<float4*>(p.grads + OFF_LSTM + e * LSTM_BLK + LSTM_BIH + 4 * unit) = z;
      *reinterpret_cast<float4*>(p.grads + OFF_LSTM + e * LSTM_BLK + LSTM_BHH + 4 * unit) = z;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 32 * LSB_NBUF);
}

}  // namespace cadre

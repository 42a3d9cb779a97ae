// Persistent, software-pipelined variant of the tcgen05 tile kernel for the encoder (fp16 in / fp16 out):
//   * one CTA per SM walks output tiles (m fastest, so a weight tile stays hot in L2);
//   * warp 0 = TMA producer running up to STAGES k-blocks ahead ACROSS tiles,
//     warp 1 = tcgen05.mma issuer alternating between two TMEM accumulators,
//     warps 2..5 = epilogue: residual row prefetched before the accumulator is ready, tcgen05.ld, bias /
//     residual / activation, fp16 pack into a 128B-swizzled staging tile, one TMA store per 64 channels;
//   * the epilogue of tile i overlaps the main loop of tile i+1 (double-buffered TMEM, mbarrier hand-off).
// Modes as in tc_gemm.cuh (GEMM / implicit-GEMM conv / stem).
#pragma once
#include "tc_gemm.cuh"

namespace cadre {

struct PersistParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB;
  CUtensorMap tmOut;
  int num_kb;
  int kb_main;              // k-blocks of the filter taps; k-blocks beyond belong to tap `ntaps` (fused shortcut)
  int ntaps, cin_chunks, Hout, Wout, TH, TN, Bimg;
  ConvTap taps[12];
  int tiles_m, tiles_n;
  int M, N;                 // logical output rows (GEMM mode) / channels
  const float* bias;
  const enc_t* res;         // residual, same layout as the output (row stride ldr elements)
  long long ldr;
  int res_after_act, act;
  int out_pad;              // output (and residual) tensors carry a 1-pixel zero border: [B][H+2][W+2][C]
  long long* dbg;           // optional per-CTA cycle counters [16] (cadre_debug_clk), nullptr in production
};

template <int BLOCK_N, int STAGES, bool CTA2 = false>
struct PersistSmem {
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = (CTA2 ? BLOCK_N / 2 : BLOCK_N) * 128;  // a CTA pair splits the B tile
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OUT_BYTES = (BLOCK_N / 64) * 128 * 128;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + OUT_BYTES + (2 * STAGES + 4) * 8 + 16 + 1024;
};

__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* src, int c0, int c1, int c2,
                                             int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read1() {
  asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync256() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// ---- cta_group::2 (CTA pair) helpers. Shared::cluster addresses carry the CTA rank in bit 24; clearing it
// addresses the same offset in the leader (even) CTA (cute/arch/copy_sm100_tma.hpp Sm100MmaPeerBitMask).
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* m, uint64_t* leader_bar, int c0,
                                                int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d_2sm(void* dst, const CUtensorMap* m, uint64_t* leader_bar, int c0,
                                                int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(leader_bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2),
      "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tc_mma_f16_2sm(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at the same offset in BOTH CTAs of the pair once the issued MMAs have completed
__device__ __forceinline__ void tc_commit_2sm(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {  // arrive on the leader CTA's barrier
  asm volatile(
      "{\n\t"
      ".reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, 0;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t"
      "}" ::"r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_2sm() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// CTA2 = true: the kernel is launched in clusters of two CTAs that compute one 256 x BLOCK_N tile with
// tcgen05.mma.cta_group::2 (M = 256): each CTA loads its own 128 rows of A and HALF of the B tile, the leader
// CTA issues the MMAs for both, each CTA drains the 128 accumulator rows in its own TMEM. Per FLOP this halves
// the B-operand shared-memory traffic (TMA write + UMMA read), which is what bounds the 1-CTA kernel.
template <int BLOCK_N, int STAGES, int MODE, bool CTA2 = false>
__global__ void __launch_bounds__(320, 1) tc_persist_kernel(const __grid_constant__ PersistParams p) {
  pdl_trigger();
  constexpr int BK = 64, UMMA_K = 16, NCH = BLOCK_N / 64;
  using S = PersistSmem<BLOCK_N, STAGES, CTA2>;
  static_assert(BLOCK_N == 64 || BLOCK_N == 128 || BLOCK_N == 256, "BLOCK_N");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* out_s = smem + STAGES * S::STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(out_s + S::OUT_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;   // [2]
  uint64_t* tempty = tfull + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = CTA2 ? static_cast<int>(cluster_ctarank()) : 0;   // 0 = leader of the pair
  const int tiles_m_sched = CTA2 ? (p.tiles_m + 1) / 2 : p.tiles_m;     // pairs of 128-row m-tiles
  const int total_tiles = tiles_m_sched * p.tiles_n;
  const int worker = CTA2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int num_workers = CTA2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  constexpr uint32_t TMEM_COLS = 2 * BLOCK_N;  // 128 / 256 / 512 columns

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmB);
    tma_prefetch_desc(&p.tmOut);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull[a], 1);
      mbar_init(&tempty[a], CTA2 ? 16 : 8);  // one arrival per epilogue warp (of both CTAs)
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (CTA2) {
      tmem_alloc_2sm(tmem_slot, TMEM_COLS);
      tmem_relinquish_2sm();
    } else {
      tmem_alloc(tmem_slot, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if constexpr (CTA2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // nothing above reads or writes global memory

  auto tile_origin = [&](int tile, int& m_tile, int& n0, int& img0, int& h0) {
    const int n_tile = tile / tiles_m_sched;
    m_tile = tile - n_tile * tiles_m_sched;
    if constexpr (CTA2) m_tile = 2 * m_tile + rank;  // this CTA's half of the 256-row tile
    n0 = n_tile * BLOCK_N;
    img0 = 0, h0 = 0;
    if constexpr (MODE != MODE_GEMM) {
      const int tiles_h = p.Hout / p.TH;
      img0 = (m_tile / tiles_h) * p.TN;
      h0 = (m_tile % tiles_h) * p.TH;
    }
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (converged warp, elected lane)
    int it = 0;  // global k-block counter across tiles
    for (int tile = worker; tile < total_tiles; tile += num_workers) {
      int m_tile, n0, img0, h0;
      tile_origin(tile, m_tile, n0, img0, h0);
      for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        const long long t0 = p.dbg ? clock64() : 0;
        mbar_wait(&empty[s], ph ^ 1);
        if (p.dbg && lane == 0) p.dbg[blockIdx.x * 16 + 1] += clock64() - t0;
        uint8_t* a_s = smem + s * S::STAGE_BYTES;
        uint8_t* b_s = a_s + S::A_BYTES;
        if (elect_one()) {
        if constexpr (CTA2) {
          // both CTAs load into their own smem; all bytes are accounted on the LEADER's full barrier
          if (rank == 0) mbar_expect_tx(&full[s], 2 * S::STAGE_BYTES);
          if constexpr (MODE == MODE_GEMM) {
            tma_load_3d_2sm(a_s, &p.tmA[0], &full[s], kb * BK, m_tile * 128, 0);
          } else if constexpr (MODE == MODE_CONV) {
            int tap = kb / p.cin_chunks, cc = kb - tap * p.cin_chunks;
            if (kb >= p.kb_main) tap = p.ntaps, cc = kb - p.kb_main;   // fused 1x1 / stride-2 shortcut
            const ConvTap t = p.taps[tap];
            tma_load_4d_2sm(a_s, &p.tmA[t.map], &full[s], cc * 64, t.dw, h0 + t.dh, img0);
          } else {
            tma_load_4d_2sm(a_s, &p.tmA[0], &full[s], 0, 0, h0 + kb, img0);
          }
          tma_load_3d_2sm(b_s, &p.tmB, &full[s], kb * BK, n0 + rank * (BLOCK_N / 2), 0);
        } else {
          mbar_expect_tx(&full[s], S::STAGE_BYTES);
          if constexpr (MODE == MODE_GEMM) {
            tma_load_3d(a_s, &p.tmA[0], &full[s], kb * BK, m_tile * 128, 0);
          } else if constexpr (MODE == MODE_CONV) {
            int tap = kb / p.cin_chunks, cc = kb - tap * p.cin_chunks;
            if (kb >= p.kb_main) tap = p.ntaps, cc = kb - p.kb_main;   // fused 1x1 / stride-2 shortcut
            const ConvTap t = p.taps[tap];
            tma_load_4d(a_s, &p.tmA[t.map], &full[s], cc * 64, t.dw, h0 + t.dh, img0);
          } else {
            tma_load_4d(a_s, &p.tmA[0], &full[s], 0, 0, h0 + kb, img0);
          }
          tma_load_3d(b_s, &p.tmB, &full[s], kb * BK, n0, 0);
        }
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
    // ------------------------------------------------------------ MMA issuer (leader CTA of a pair;
    // converged warp, elected lane)
    constexpr uint32_t idesc = umma_idesc(CADRE_ENC_FP16 ? 0u : 1u, 0, 0, CTA2 ? 256 : 128, BLOCK_N);
    int it = 0, lt = 0;
    long long d_te = 0, d_wf = 0;
    const long long tstart = p.dbg ? clock64() : 0;
    for (int tile = worker; tile < total_tiles; tile += num_workers, ++lt) {
      const int as = lt & 1;
      const uint32_t aph = (lt >> 1) & 1;
      const long long t0 = p.dbg ? clock64() : 0;
      mbar_wait(&tempty[as], aph ^ 1);  // epilogue has drained this accumulator
      if (p.dbg) d_te += clock64() - t0;
      tc_fence_after();
      const uint32_t tacc = tmem_base + as * BLOCK_N;
      for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (it / STAGES) & 1;
        const long long t1 = p.dbg ? clock64() : 0;
        mbar_wait(&full[s], ph);
        if (p.dbg) d_wf += clock64() - t1;
        tc_fence_after();
        const uint32_t a_addr = smem_u32(smem + s * S::STAGE_BYTES);
        const uint32_t b_addr = a_addr + S::A_BYTES;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = umma_smem_desc(a_addr + k * 32, 16, 1024, 2);
            const uint64_t db = umma_smem_desc(b_addr + k * 32, 16, 1024, 2);
            if constexpr (CTA2)
              tc_mma_f16_2sm(tacc, da, db, idesc, (kb | k) != 0);
            else
              tc_mma_f16(tacc, da, db, idesc, (kb | k) != 0);
          }
          if constexpr (CTA2) tc_commit_2sm(&empty[s]); else tc_commit(&empty[s]);
          if (kb == p.num_kb - 1) {
            if constexpr (CTA2) tc_commit_2sm(&tfull[as]); else tc_commit(&tfull[as]);
          }
        }
        __syncwarp();
      }
    }
    if (p.dbg && lane == 0) {
      long long* d = p.dbg + blockIdx.x * 16;
      d[2] += d_te, d[3] += d_wf, d[5] += clock64() - tstart, d[10] += lt;
    }
    }
  } else if (warp >= 2) {
    // ------------------------------------------------------------ epilogue: 8 warps, 2 per TMEM lane quarter;
    // warps 2..5 own the first half of the tile's columns, warps 6..9 the second half
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = q * 32 + lane;
    const bool leader = (warp == 2 && lane == 0);
    constexpr int HC = BLOCK_N / 2;          // columns per thread
    const int cbase = half * HC;             // first column (within the tile) of this thread
    int lt = 0;
    for (int tile = worker; tile < total_tiles; tile += num_workers, ++lt) {
      int m_tile, n0, img0, h0;
      tile_origin(tile, m_tile, n0, img0, h0);
      const int as = lt & 1;
      const uint32_t aph = (lt >> 1) & 1;
      // row -> residual row index (same layout as the output)
      long long out_row;
      bool row_ok;
      if constexpr (MODE == MODE_GEMM) {
        out_row = static_cast<long long>(m_tile) * 128 + row;
        row_ok = out_row < p.M;
      } else {
        const int w = row % p.Wout;
        const int hh = (row / p.Wout) % p.TH;
        const int im = row / (p.Wout * p.TH);
        const int img = img0 + im, h = h0 + hh;
        row_ok = img < p.Bimg;
        if (p.out_pad)
          out_row = (static_cast<long long>(img) * (p.Hout + 2) + h + 1) * (p.Wout + 2) + w + 1;
        else
          out_row = (static_cast<long long>(img) * p.Hout + h) * p.Wout + w;
      }
      // prefetch this thread's residual half-row while the MMAs of the tile are still in flight
      uint4 rres[HC / 8];
      const bool has_res = p.res != nullptr;
      const bool relu_at_pack = p.act == ACT_RELU && !(has_res && p.res_after_act);
      if (has_res) {
        const uint4* rp = reinterpret_cast<const uint4*>(p.res + out_row * p.ldr + n0 + cbase);
#pragma unroll
        for (int j = 0; j < HC / 8; ++j) rres[j] = row_ok ? __ldg(rp + j) : make_uint4(0, 0, 0, 0);
      }
      // staging buffer must have been read by the previous tile's TMA store
      const long long e0 = (p.dbg && leader) ? clock64() : 0;
      if (leader) tma_store_wait_read();
      epi_bar_sync256();
      const long long e1 = (p.dbg && leader) ? clock64() : 0;
      mbar_wait(&tfull[as], aph);
      const long long e2 = (p.dbg && leader) ? clock64() : 0;
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * BLOCK_N + cbase + (static_cast<uint32_t>(q * 32) << 16);
#pragma unroll
      for (int c = 0; c < HC / 32; ++c) {
        uint32_t r[32];
        const long long l0 = (p.dbg && leader) ? clock64() : 0;
        tmem_ld_32x32(taddr + c * 32, r);
        tmem_ld_wait();
        if (p.dbg && leader) p.dbg[blockIdx.x * 16 + 11] += clock64() - l0;
        if (c == HC / 32 - 1) {  // this warp's part of the accumulator is in registers
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if constexpr (CTA2) mbar_arrive_leader(&tempty[as]); else mbar_arrive(&tempty[as]);
          }
        }
        const int col = cbase + c * 32;  // column within the tile
        const int nb = n0 + col;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float v[8];
          float4 b_lo = make_float4(0.f, 0.f, 0.f, 0.f), b_hi = b_lo;
          if (nb + 8 * j < p.N) {  // N is a multiple of 8 on this path: two 16-byte bias loads per 8 columns
            b_lo = __ldg(reinterpret_cast<const float4*>(p.bias + nb + 8 * j));
            b_hi = __ldg(reinterpret_cast<const float4*>(p.bias + nb + 8 * j + 4));
          }
          const float bb[8] = {b_lo.x, b_lo.y, b_lo.z, b_lo.w, b_hi.x, b_hi.y, b_hi.z, b_hi.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[8 * j + i]) + bb[i];
          if (has_res && !p.res_after_act) {
            const enc_t* h8 = reinterpret_cast<const enc_t*>(&rres[c * 4 + j]);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += enc_to_float(h8[i]);
          }
          if (p.act == ACT_RELU && !relu_at_pack) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = fmaxf(v[i], 0.f);
          } else if (p.act == ACT_LEAKY) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = v[i] > 0.f ? v[i] : 0.01f * v[i];
          }
          if (has_res && p.res_after_act) {
            const enc_t* h8 = reinterpret_cast<const enc_t*>(&rres[c * 4 + j]);
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] += enc_to_float(h8[i]);
          }
          // 16-byte chunk of this row inside 64-channel group g; 128B swizzle: chunk ^= row & 7
          const int g = (col + 8 * j) >> 6;
          const int chunk = ((col + 8 * j) & 63) >> 3;
          uint4 u;
          if (relu_at_pack) {   // ReLU folded into the fp16 conversion
            u.x = enc_pack2_relu(v[0], v[1]), u.y = enc_pack2_relu(v[2], v[3]);
            u.z = enc_pack2_relu(v[4], v[5]), u.w = enc_pack2_relu(v[6], v[7]);
          } else {
            u.x = enc_pack2(v[0], v[1]), u.y = enc_pack2(v[2], v[3]);
            u.z = enc_pack2(v[4], v[5]), u.w = enc_pack2(v[6], v[7]);
          }
          *reinterpret_cast<uint4*>(out_s + g * (128 * 128) + row * 128 + ((chunk ^ (row & 7)) << 4)) = u;
        }
      }
      const long long e3 = (p.dbg && leader) ? clock64() : 0;
      fence_proxy_async_smem();
      epi_bar_sync256();
      if (p.dbg && leader) p.dbg[blockIdx.x * 16 + 12] += clock64() - e3;
      if (leader) {
#pragma unroll
        for (int g = 0; g < NCH; ++g) {
          if (n0 + g * 64 < p.N) {
            if constexpr (MODE == MODE_GEMM)
              tma_store_3d(&p.tmOut, out_s + g * (128 * 128), n0 + g * 64, m_tile * 128, 0);
            else
              tma_store_4d(&p.tmOut, out_s + g * (128 * 128), n0 + g * 64, p.out_pad, h0 + p.out_pad, img0);
          }
        }
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
  if constexpr (CTA2) {
    cluster_sync_all();  // the peer may still multicast into this CTA's barriers / read its smem
    if (warp == 1) tmem_dealloc_2sm(tmem_base, TMEM_COLS);
  } else {
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

}  // namespace cadre

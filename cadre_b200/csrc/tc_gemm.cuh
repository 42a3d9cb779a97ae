// tcgen05 / TMEM / TMA tile kernel shared by every dense contraction on the CADRE learner path:
//   * MODE_GEMM  plain (optionally grid.z-batched) GEMM  D[M,N] = A[M,K] * B[N,K]^T  with either operand
//                K-major or MN-major in global memory (covers fwd, dgrad and wgrad without transposes);
//   * MODE_CONV  implicit-GEMM convolution over NHWC activations: one TMA box per filter tap, padding by
//                TMA out-of-bounds zero fill, stride 2 through parity sub-lattice tensor maps;
//   * MODE_STEM  the 7x7/s2 Cin=4 stem through an overlapping-stride 5-D tensor map.
// One CTA = one 128 x BLOCK_N output tile. Warp 0 lane 0 issues TMA, warp 1 lane 0 issues tcgen05.mma into a
// TMEM accumulator, warps 2..5 drain TMEM (tcgen05.ld) and run the fused epilogue.
#pragma once
#include "internal.h"
#include "ptx.cuh"

namespace cadre {

enum { MODE_GEMM = 0, MODE_CONV = 1, MODE_STEM = 2 };
enum { EPI_LINEAR = 0, EPI_LSTM = 1 };
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_LEAKY = 2 };

struct ConvTap {
  short map, dw, dh, pad_;
};

struct TcGemmParams {
  CUtensorMap tmA[4];
  CUtensorMap tmB;
  int M, N, num_kb;
  // conv geometry (output map); TW == Wout
  int ntaps, cin_chunks, Hout, Wout, TH, TN, Bimg;
  ConvTap taps[12];
  // epilogue
  void* out;
  long long ldc, out_bs;
  const float* bias;
  long long bias_bs;
  const void* res;  // same dtype as out
  long long ldr, res_bs;
  int res_after_act;
  const void* mask;  // same dtype as out; output zeroed where mask <= 0 (ReLU backward)
  long long ldm, mask_bs;
  int act;
  float alpha;
  int dbg_a_shift, dbg_base_offset;  // experiment: A descriptor start shifted by rows (128 B each)
  const int* batch_rows;  // optional [gridDim.z]: valid rows (M) per batch, or valid K when rows_is_k
  int rows_is_k;
  // EPI_LSTM: acc = h_{t-1} W_hh^T (gate-interleaved columns 4*u+g); xpart holds x_t W_ih^T + b_ih + b_hh
  const float* xpart;
  long long ldx, x_bs;
  const float* c_prev;
  float* c_out;
  float* h_out;
  float* gates_out;
  long long ldh, h_bs;
};

template <int KIND, int BLOCK_N, int STAGES>
struct TcGemmSmem {
  static constexpr int A_BYTES = 128 * 128;
  static constexpr int B_BYTES = BLOCK_N * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TOTAL = STAGES * STAGE_BYTES + (2 * STAGES + 1) * 8 + 16 + 1024;
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

template <typename OutT>
__device__ __forceinline__ float load_as_float(const void* base, long long idx) {
  if constexpr (sizeof(OutT) == 2)
    return enc_to_float(reinterpret_cast<const enc_t*>(base)[idx]);
  else
    return reinterpret_cast<const float*>(base)[idx];
}

template <int KIND, int A_MN, int B_MN, int BLOCK_N, int STAGES, int MODE, int EPI, typename OutT>
__global__ void __launch_bounds__(192) tc_gemm_kernel(const __grid_constant__ TcGemmParams p) {
  constexpr int ES = KIND ? 4 : 2;         // operand element bytes
  constexpr int BK = 128 / ES;             // K elements per stage (one 128-byte swizzle row)
  constexpr int CHUNK = 128 / ES;          // MN elements per 128-byte row of an MN-major operand
  constexpr int UMMA_K = 32 / ES;          // K per tcgen05.mma
  using S = TcGemmSmem<KIND, BLOCK_N, STAGES>;
  static_assert(BLOCK_N >= 32 && BLOCK_N <= 256 && (BLOCK_N & (BLOCK_N - 1)) == 0, "BLOCK_N");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * S::STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tmem_full = empty + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tile = blockIdx.x, n_tile = blockIdx.y, batch = blockIdx.z;
  const int n0 = n_tile * BLOCK_N;

  int m_valid = p.M;
  int num_kb = p.num_kb;
  if (p.batch_rows != nullptr) {
    const int cnt = p.batch_rows[batch];
    if (p.rows_is_k) {
      num_kb = (cnt + BK - 1) / BK;  // operands are zero padded up to the next multiple of 128 rows
    } else {
      m_valid = cnt;
      if (m_tile * 128 >= cnt) return;  // uniform for the whole CTA, before any barrier
    }
  }

  // conv tile origin
  int img0 = 0, h0 = 0;
  if constexpr (MODE != MODE_GEMM) {
    const int tiles_h = p.Hout / p.TH;
    img0 = (m_tile / tiles_h) * p.TN;
    h0 = (m_tile % tiles_h) * p.TH;
  }

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tmA[0]);
    tma_prefetch_desc(&p.tmB);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, BLOCK_N);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ------------------------------------------------------------ TMA producer
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&empty[s], ph ^ 1);
      mbar_expect_tx(&full[s], S::STAGE_BYTES);
      uint8_t* a_s = smem + s * S::STAGE_BYTES;
      uint8_t* b_s = a_s + S::A_BYTES;
      if constexpr (MODE == MODE_GEMM) {
        if constexpr (A_MN) {
#pragma unroll
          for (int c = 0; c < 128 / CHUNK; ++c)
            tma_load_3d(a_s + c * BK * 128, &p.tmA[0], &full[s], m_tile * 128 + c * CHUNK, kb * BK, batch);
        } else {
          tma_load_3d(a_s, &p.tmA[0], &full[s], kb * BK, m_tile * 128, batch);
        }
      } else if constexpr (MODE == MODE_CONV) {
        const int tap = kb / p.cin_chunks, cc = kb - tap * p.cin_chunks;
        const ConvTap t = p.taps[tap];
        tma_load_4d(a_s, &p.tmA[t.map], &full[s], cc * 64, t.dw, h0 + t.dh, img0);
      } else {  // MODE_STEM: k-block kb covers filter rows 2kb, 2kb+1 (x 8 pixels x 4 channels)
        tma_load_4d(a_s, &p.tmA[0], &full[s], 0, 0, h0 + kb, img0);
      }
      if constexpr (B_MN) {
#pragma unroll
        for (int c = 0; c < BLOCK_N / CHUNK; ++c)
          tma_load_3d(b_s + c * BK * 128, &p.tmB, &full[s], n0 + c * CHUNK, kb * BK, batch);
      } else {
        tma_load_3d(b_s, &p.tmB, &full[s], kb * BK, n0, batch);
      }
    }
  } else if (warp == 1 && lane == 0) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = umma_idesc(KIND ? 2u : (CADRE_ENC_FP16 ? 0u : 1u), A_MN, B_MN, 128, BLOCK_N);
    constexpr uint32_t A_KSTEP = A_MN ? UMMA_K * 128 : 32;
    constexpr uint32_t B_KSTEP = B_MN ? UMMA_K * 128 : 32;
    constexpr uint32_t A_LBO = A_MN ? BK * 128 : 16;
    constexpr uint32_t B_LBO = B_MN ? BK * 128 : 16;
    // MN-major TF32 uses SWIZZLE_128B_BASE32B (layout type 1): atoms of 4 K-rows x 128 B -> SBO = 512 B
    constexpr uint32_t A_SBO = (A_MN && KIND) ? 512 : 1024, B_SBO = (B_MN && KIND) ? 512 : 1024;
    constexpr uint32_t A_LAYOUT = (A_MN && KIND) ? 1 : 2, B_LAYOUT = (B_MN && KIND) ? 1 : 2;
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(&full[s], ph);
      tc_fence_after();
      const uint32_t a_addr = smem_u32(smem + s * S::STAGE_BYTES);
      const uint32_t b_addr = a_addr + S::A_BYTES;
#pragma unroll
      for (int k = 0; k < BK / UMMA_K; ++k) {
        uint64_t da = umma_smem_desc(a_addr + k * A_KSTEP + p.dbg_a_shift * 128, A_LBO, A_SBO, A_LAYOUT);
        da |= static_cast<uint64_t>(p.dbg_base_offset & 7) << 49;
        const uint64_t db = umma_smem_desc(b_addr + k * B_KSTEP, B_LBO, B_SBO, B_LAYOUT);
        if constexpr (KIND)
          tc_mma_tf32(tmem_base, da, db, idesc, (kb | k) != 0);
        else
          tc_mma_f16(tmem_base, da, db, idesc, (kb | k) != 0);
      }
      tc_commit(&empty[s]);  // frees the smem stage once these MMAs have read it
    }
    tc_commit(tmem_full);
  } else if (warp >= 2) {
    // ------------------------------------------------------------ epilogue (TMEM -> registers -> global)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;
    long long out_row;  // row index into out/res/mask (units of rows)
    bool row_ok;
    if constexpr (MODE == MODE_GEMM) {
      const int m = m_tile * 128 + row;
      out_row = m;
      row_ok = m < m_valid;
    } else {
      const int w = row % p.Wout;
      const int hh = (row / p.Wout) % p.TH;
      const int im = row / (p.Wout * p.TH);
      const int img = img0 + im, h = h0 + hh;
      row_ok = img < p.Bimg && im < p.TN;
      out_row = (static_cast<long long>(img) * p.Hout + h) * p.Wout + w;
    }
    // rows of a batched operand beyond batch_rows are written as zeros (keeps K-padding of later GEMMs clean)
    const bool zero_fill = (p.batch_rows != nullptr) && !p.rows_is_k && !row_ok &&
                           (MODE == MODE_GEMM) && (m_tile * 128 + row) < p.M;
    if (num_kb > 0) {
      mbar_wait(tmem_full, 0);
      tc_fence_after();
    }
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

    if constexpr (EPI == EPI_LINEAR) {
      OutT* out = reinterpret_cast<OutT*>(p.out) + batch * p.out_bs;
      const float* bias = p.bias ? p.bias + batch * p.bias_bs : nullptr;
      const OutT* res = p.res ? reinterpret_cast<const OutT*>(p.res) + batch * p.res_bs : nullptr;
      const OutT* mask = p.mask ? reinterpret_cast<const OutT*>(p.mask) + batch * p.mask_bs : nullptr;
      const bool vec_ok = (p.ldc * sizeof(OutT)) % 16 == 0 && (reinterpret_cast<uintptr_t>(out) % 16 == 0);
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t r[32];
        if (num_kb > 0) {
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = 0u;
        }
        const int nb = n0 + c * 32;
        if (nb < p.N && (row_ok || zero_fill)) {
        float v[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float x = __uint_as_float(r[i]) * p.alpha;
          const int n = nb + i;
          if (n < p.N && row_ok) {
            if (bias) x += bias[n];
            if (res && !p.res_after_act) x += load_as_float<OutT>(res, out_row * p.ldr + n);
            if (p.act == ACT_RELU) x = fmaxf(x, 0.f);
            if (p.act == ACT_LEAKY) x = x > 0.f ? x : 0.01f * x;
            if (res && p.res_after_act) x += load_as_float<OutT>(res, out_row * p.ldr + n);
            if (mask && !(load_as_float<OutT>(mask, out_row * p.ldm + n) > 0.f)) x = 0.f;
          } else {
            x = 0.f;
          }
          v[i] = x;
        }
        OutT* dst = out + out_row * p.ldc + nb;
        if (vec_ok && nb + 32 <= p.N) {
          if constexpr (sizeof(OutT) == 2) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 u;
              u.x = enc_pack2(v[8 * i + 0], v[8 * i + 1]);
              u.y = enc_pack2(v[8 * i + 2], v[8 * i + 3]);
              u.z = enc_pack2(v[8 * i + 4], v[8 * i + 5]);
              u.w = enc_pack2(v[8 * i + 6], v[8 * i + 7]);
              reinterpret_cast<uint4*>(dst)[i] = u;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              reinterpret_cast<float4*>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (nb + i < p.N) {
              if constexpr (sizeof(OutT) == 2)
                dst[i] = enc_from_float(v[i]);
              else
                dst[i] = v[i];
            }
        }
        }
        __syncwarp();
      }
    } else {
      // EPI_LSTM: nn.LSTMCell pointwise part (ppo_agent/models.py:139-152 -> torch LSTMCell, gate order
      // i,f,g,o); columns are gate-interleaved so one thread holds all four gates of 8 hidden units.
      const float* xpart = p.xpart + batch * p.x_bs;
      const float* c_prev = p.c_prev + batch * p.h_bs;
      float* c_out = p.c_out + batch * p.h_bs;
      float* h_out = p.h_out + batch * p.h_bs;
      float* gates_out = p.gates_out + batch * p.x_bs;
#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t r[32];
        if (num_kb > 0) {
          tmem_ld_32x32(taddr + c * 32, r);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) r[i] = 0u;
        }
        const int nb = n0 + c * 32;
        if (nb < p.N && (row_ok || zero_fill)) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int n = nb + 4 * j;
          if (n >= p.N) continue;
          const int u = n >> 2;
          float gi = 0.f, gf = 0.f, gg = 0.f, go = 0.f, cn = 0.f, hn = 0.f;
          if (row_ok) {
            const float4 xp = *reinterpret_cast<const float4*>(xpart + out_row * p.ldx + n);
            gi = sigmoidf_(__uint_as_float(r[4 * j + 0]) + xp.x);
            gf = sigmoidf_(__uint_as_float(r[4 * j + 1]) + xp.y);
            gg = tanhf(__uint_as_float(r[4 * j + 2]) + xp.z);
            go = sigmoidf_(__uint_as_float(r[4 * j + 3]) + xp.w);
            cn = gf * c_prev[out_row * p.ldh + u] + gi * gg;
            hn = go * tanhf(cn);
          }
          *reinterpret_cast<float4*>(gates_out + out_row * p.ldx + n) = make_float4(gi, gf, gg, go);
          c_out[out_row * p.ldh + u] = cn;
          h_out[out_row * p.ldh + u] = hn;
        }
        }
        __syncwarp();
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, BLOCK_N);
}

}  // namespace cadre

// Non-GEMM kernels of the perception encoder forward pass (HBM / latency bound, fp32 math on CUDA cores):
// uint8 ingest, max-pool, fused position attention (PAM), fused channel attention (CAM), inter-task attention.
#include "internal.h"
#include "ptx.cuh"
#include <stdlib.h>

namespace cadre {

// ---------------------------------------------------------------------------------------------------------
// pre_process (ppo_agent/agent.py:43-75): rgb u8 [B,144,256,3] / 255 -> channels 0..2; route_fig u8
// [B,256,144], max-normalised per frame and truncated back to uint8 (so 1 where route == max > 0, else 0),
// transposed -> channel 3. Output: the stem's row-pair interleaved, 3-pixel padded bf16 image
// P[B][75][262][2][4] (borders stay zero from allocation time).
__global__ void route_max_kernel(const uint8_t* __restrict__ route, uint8_t* __restrict__ mx) {
  pdl_trigger();
  pdl_wait();
  const uint8_t* r = route + static_cast<long long>(blockIdx.x) * 256 * 144;
  unsigned m = 0;
  const uint4* r4 = reinterpret_cast<const uint4*>(r);
  for (int i = threadIdx.x; i < 256 * 144 / 16; i += blockDim.x) {
    uint4 v = r4[i];
    unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      m = max(m, w[k] & 0xff);
      m = max(m, (w[k] >> 8) & 0xff);
      m = max(m, (w[k] >> 16) & 0xff);
      m = max(m, w[k] >> 24);
    }
  }
  __shared__ unsigned sm[32];
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
  __syncthreads();
  if (threadIdx.x < 32) {
    m = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 0;
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (threadIdx.x == 0) mx[blockIdx.x] = static_cast<uint8_t>(m);
  }
}

// One block = 4 x 32 pixels x 8 row pairs of one image. The route map is [n][x][y] (y contiguous) while the output is
// x-major, so its 32 x 16 byte tile is transposed through shared memory (lanes along y when loading: whole
// sectors; the direct x-strided byte reads fetched every sector 16 times). k / 255 comes from a 256-entry table
// (same double division + float + fp16 roundings as the reference, done once on the host instead of three
// FP64 divisions per pixel).
__global__ void __launch_bounds__(256) preprocess_kernel(const uint8_t* __restrict__ rgb,
                                                         const uint8_t* __restrict__ route,
                                                         const uint8_t* __restrict__ route_max,
                                                         const enc_t* __restrict__ lut_g,
                                                         enc_t* __restrict__ out, int B) {
  pdl_trigger();
  pdl_wait();
  constexpr int XT = 4;                     // 32-pixel tiles per block (fewer, fatter blocks: block launch rate
                                            // bounded the one-tile version)
  __shared__ enc_t lut[256];
  __shared__ uint8_t s_route[32 * XT][20];  // [x][y - ybase], 16 rows used
  const int tid = threadIdx.x;
  const int x0 = blockIdx.x * 32 * XT;      // 256 / (32 * XT) blocks across
  const int yb = blockIdx.y;                // 0..9: row pairs y2 = 1 + 8*yb .. (73 row pairs: the last block has one)
  const int n = blockIdx.z;
  const int ybase = 2 * (1 + 8 * yb) - 3;   // unpadded row of (y2 = first, r = 0)
  const int lx = tid & 31, ly = tid >> 5;   // pixel, row pair within the tile
  const int y2 = 1 + 8 * yb + ly;
  // every global load of the block is issued before the barrier: one memory round trip per block
  const enc_t lut_v = lut_g[tid];
  uint8_t rt[XT][2];
  {
    const int ty = tid & 15, tx = tid >> 4;   // 16 lanes along y, 16 x rows per pass
    const int y = ybase + ty;
    const bool yok = y >= 0 && y < 144;
#pragma unroll
    for (int k = 0; k < 2 * XT; ++k)
      rt[k >> 1][k & 1] = yok ? route[(static_cast<long long>(n) * 256 + x0 + tx + 16 * k) * 144 + y] : 0;
  }
  uint8_t pb[XT][2][3];
#pragma unroll
  for (int xt = 0; xt < XT; ++xt) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int y = 2 * y2 + r - 3;  // unpadded row
      const bool ok = y2 <= 73 && y >= 0 && y < 144;
      const uint8_t* px = rgb + ((static_cast<long long>(n) * 144 + (ok ? y : 0)) * 256 + x0 + 32 * xt + lx) * 3;
      pb[xt][r][0] = ok ? px[0] : 0, pb[xt][r][1] = ok ? px[1] : 0, pb[xt][r][2] = ok ? px[2] : 0;
    }
  }
  const uint8_t mx = route_max[n];
  lut[tid] = lut_v;   // k / 255 as the reference rounds it (table built once on the host)
#pragma unroll
  for (int k = 0; k < 2 * XT; ++k) s_route[(tid >> 4) + 16 * k][tid & 15] = rt[k >> 1][k & 1];
  __syncthreads();
  if (y2 > 73) return;
#pragma unroll
  for (int xt = 0; xt < XT; ++xt) {
    enc_t v[8];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int y = 2 * y2 + r - 3;  // unpadded row
      if (y >= 0 && y < 144) {
        v[4 * r + 0] = lut[pb[xt][r][0]];
        v[4 * r + 1] = lut[pb[xt][r][1]];
        v[4 * r + 2] = lut[pb[xt][r][2]];
        const uint8_t rv = s_route[32 * xt + lx][y - ybase];
        v[4 * r + 3] = enc_from_float((mx > 0) ? ((rv == mx) ? 1.f : 0.f) : static_cast<float>(rv));
      } else {
        v[4 * r + 0] = v[4 * r + 1] = v[4 * r + 2] = v[4 * r + 3] = enc_from_float(0.f);
      }
    }
    enc_t* dst = out + ((static_cast<long long>(n) * 75 + y2) * 262 + (x0 + 32 * xt + lx + 3)) * 8;
    *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(v);
  }
}

__global__ void f32_to_enc_kernel(const float* __restrict__ in, enc_t* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = enc_from_float(in[i]);   // same rounding as every other fp32 -> operand conversion
}
void launch_f32_to_enc(const float* in, enc_t* out, int n, cudaStream_t stream) {
  f32_to_enc_kernel<<<(n + 255) / 256, 256, 0, stream>>>(in, out, n);
  CADRE_CUDA_CHECK(cudaGetLastError());
}

void launch_preprocess(const uint8_t* rgb, const uint8_t* route, uint8_t* route_max_ws, const enc_t* lut,
                       enc_t* out, int B, cudaStream_t stream) {
  launch_k(route_max_kernel, dim3(B), dim3(256), 0, stream, route, route_max_ws);
  launch_k(preprocess_kernel, dim3(2, 10, B), dim3(256), 0, stream, rgb, route, route_max_ws, lut, out, B);
  CADRE_CUDA_CHECK(cudaGetLastError());
}

// same layout from an already-normalised fp32 NCHW tensor [B,4,144,256] (parity tests / encoder sweep input)
__global__ void pack_f32_kernel(const float* __restrict__ x, enc_t* __restrict__ out, int B) {
  pdl_trigger();
  pdl_wait();
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long total = static_cast<long long>(B) * 73 * 256;
  if (gid >= total) return;
  const int xx = static_cast<int>(gid % 256);
  const int y2 = static_cast<int>((gid / 256) % 73) + 1;
  const int n = static_cast<int>(gid / (256 * 73));
  enc_t v[8];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int y = 2 * y2 + r - 3;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      float f = 0.f;
      if (y >= 0 && y < 144) f = x[((static_cast<long long>(n) * 4 + c) * 144 + y) * 256 + xx];
      v[4 * r + c] = enc_from_float(f);
    }
  }
  enc_t* dst = out + ((static_cast<long long>(n) * 75 + y2) * 262 + (xx + 3)) * 8;
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(v);
}

void launch_pack_f32(const float* x, enc_t* out, int B, cudaStream_t stream) {
  const long long total = static_cast<long long>(B) * 73 * 256;
  launch_k(pack_f32_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, stream, x, out, B);
  CADRE_CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------------
// Position / channel attention (danet_blocks/da_att.py:32-83): 40 positions x 128 channels per frame.
constexpr int PAM_P = 40, PAM_C = 128;

#if CADRE_ENC_FP16
// ---------------------------------------------------------------------------------------------------------
// CAM on tensor cores. The shapes are tiny (128x128x40 gram, 128x40x128 product per frame), so the kernel uses
// warp-level mma.sync.m16n8k16 (fp16 in, fp32 accumulate) and keeps the whole chain in registers, flash-attention
// style: warp w owns channel rows [16w, 16w+16): gram row block (16 n-tiles x 3 k-steps) -> rowmax/rowmin ->
// exp -> the accumulator fragments ARE the A fragments of the second product (x X, K = 128) -> 1/rowsum ->
// staged fp32 -> gamma * out + x, coalesced. X is read once from global; both products read it from shared
// memory with ldmatrix (transposed for the gram: A[m=c][k=p] and B[k=p][n=c] are X^T blocks).
struct CamMmaSmem {
  __half x[48][136];   // X[p][c]; rows 40..47 are zero (K / N padding); 272-byte pitch: conflict-free ldmatrix
  float o[40][132];    // att X, before gamma and the residual
};

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const void* p) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(smem_u32(p)));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

__global__ void __launch_bounds__(256) cam_mma_kernel(const enc_t* __restrict__ xin, enc_t* __restrict__ out,
                                                      float gamma, int B, int ldin) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t cam_raw[];
  CamMmaSmem& s = *reinterpret_cast<CamMmaSmem*>(cam_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  if (tid < 128) *reinterpret_cast<uint4*>(&s.x[40 + (tid >> 4)][(tid & 15) * 8]) = make_uint4(0, 0, 0, 0);
  for (int f = blockIdx.x; f < B; f += gridDim.x) {
    __syncthreads();   // previous frame's readers of s.x / s.o are done
    const enc_t* xf = xin + static_cast<long long>(f) * PAM_P * ldin;
    for (int i = tid; i < PAM_P * PAM_C / 8; i += 256) {
      const int p = i >> 4, c8 = (i & 15) * 8;
      *reinterpret_cast<uint4*>(&s.x[p][c8]) = __ldg(reinterpret_cast<const uint4*>(xf + p * ldin + c8));
    }
    __syncthreads();
    // ---- gram rows [16w, 16w+16): E[a][b] = sum_p X[p][a] X[p][b]
    float e[16][4];
#pragma unroll
    for (int j = 0; j < 16; ++j) e[j][0] = e[j][1] = e[j][2] = e[j][3] = 0.f;
    const int a0 = warp * 16;
#pragma unroll
    for (int kk = 0; kk < 3; ++kk) {
      const int p0 = kk * 16;
      uint32_t af[4];
      ldsm_x4_trans(af, &s.x[p0 + (lane & 7) + ((lane >> 4) & 1) * 8][a0 + ((lane >> 3) & 1) * 8]);
#pragma unroll
      for (int jp = 0; jp < 8; ++jp) {
        uint32_t bf[4];
        ldsm_x4_trans(bf, &s.x[p0 + (lane & 7) + ((lane >> 3) & 1) * 8][jp * 16 + ((lane >> 4) & 1) * 8]);
        mma_16816(e[2 * jp], af, bf[0], bf[1]);
        mma_16816(e[2 * jp + 1], af, bf[2], bf[3]);
      }
    }
    // ---- softmax(rowmax - E) (da_att.py:76-77); this thread holds 32 values of rows g (regs 0,1) and g+8 (2,3)
    float mx[2] = {-INFINITY, -INFINITY}, mn[2] = {INFINITY, INFINITY};
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      mx[0] = fmaxf(mx[0], fmaxf(e[j][0], e[j][1])), mn[0] = fminf(mn[0], fminf(e[j][0], e[j][1]));
      mx[1] = fmaxf(mx[1], fmaxf(e[j][2], e[j][3])), mn[1] = fminf(mn[1], fminf(e[j][2], e[j][3]));
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], o));
        mn[r] = fminf(mn[r], __shfl_xor_sync(0xffffffffu, mn[r], o));
      }
    }
    // energy_new = mx - e; its row maximum is mx - mn (rounding is monotonic); p = exp(energy_new - max)
    const float m2[2] = {mx[0] - mn[0], mx[1] - mn[1]};
    float sum[2] = {0.f, 0.f};
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      e[j][0] = __expf((mx[0] - e[j][0]) - m2[0]), e[j][1] = __expf((mx[0] - e[j][1]) - m2[0]);
      e[j][2] = __expf((mx[1] - e[j][2]) - m2[1]), e[j][3] = __expf((mx[1] - e[j][3]) - m2[1]);
      sum[0] += e[j][0] + e[j][1], sum[1] += e[j][2] + e[j][3];
    }
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      sum[0] += __shfl_xor_sync(0xffffffffu, sum[0], o);
      sum[1] += __shfl_xor_sync(0xffffffffu, sum[1], o);
    }
    // ---- out[a][p] = sum_b P[a][b] X[p][b]: the C fragments of tiles (2kk, 2kk+1) are the A fragment of k-step kk
    float oacc[6][4];
#pragma unroll
    for (int j = 0; j < 6; ++j) oacc[j][0] = oacc[j][1] = oacc[j][2] = oacc[j][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
      uint32_t af[4];
      af[0] = pack_h2(e[2 * kk][0], e[2 * kk][1]), af[1] = pack_h2(e[2 * kk][2], e[2 * kk][3]);
      af[2] = pack_h2(e[2 * kk + 1][0], e[2 * kk + 1][1]), af[3] = pack_h2(e[2 * kk + 1][2], e[2 * kk + 1][3]);
#pragma unroll
      for (int jp = 0; jp < 3; ++jp) {
        uint32_t bf[4];
        ldsm_x4(bf, &s.x[jp * 16 + (lane & 7) + ((lane >> 4) & 1) * 8][kk * 16 + ((lane >> 3) & 1) * 8]);
        mma_16816(oacc[2 * jp], af, bf[0], bf[1]);
        mma_16816(oacc[2 * jp + 1], af, bf[2], bf[3]);
      }
    }
    const float inv[2] = {1.f / sum[0], 1.f / sum[1]};
#pragma unroll
    for (int j = 0; j < 5; ++j) {   // n-tile 5 is the zero padding (p = 40..47)
      const int p = 8 * j + 2 * t;
      s.o[p][a0 + g] = oacc[j][0] * inv[0], s.o[p + 1][a0 + g] = oacc[j][1] * inv[0];
      s.o[p][a0 + g + 8] = oacc[j][2] * inv[1], s.o[p + 1][a0 + g + 8] = oacc[j][3] * inv[1];
    }
    __syncthreads();
    enc_t* of = out + static_cast<long long>(f) * PAM_P * PAM_C;
    for (int i = tid; i < PAM_P * PAM_C / 8; i += 256) {
      const int p = i >> 4, c8 = (i & 15) * 8;
      const float4 o_lo = *reinterpret_cast<const float4*>(&s.o[p][c8]);
      const float4 o_hi = *reinterpret_cast<const float4*>(&s.o[p][c8 + 4]);
      const uint4 ux = *reinterpret_cast<const uint4*>(&s.x[p][c8]);
      const __half* hx = reinterpret_cast<const __half*>(&ux);
      uint4 u;
      u.x = enc_pack2(gamma * o_lo.x + __half2float(hx[0]), gamma * o_lo.y + __half2float(hx[1]));
      u.y = enc_pack2(gamma * o_lo.z + __half2float(hx[2]), gamma * o_lo.w + __half2float(hx[3]));
      u.z = enc_pack2(gamma * o_hi.x + __half2float(hx[4]), gamma * o_hi.y + __half2float(hx[5]));
      u.w = enc_pack2(gamma * o_hi.z + __half2float(hx[6]), gamma * o_hi.w + __half2float(hx[7]));
      *reinterpret_cast<uint4*>(of + p * PAM_C + c8) = u;
    }
  }
}
#endif

#if CADRE_ENC_FP16
// ---------------------------------------------------------------------------------------------------------
// PAM on tensor cores, value projection included (da_att.py:32-51). Per frame, eight warps:
//   1. [q | k | V] = X [W_q | W_k | W_v]^T + b on mma.sync.m16n8k16. The query / key weights are fp32 split into
//      fp16 (hi, lo) pairs and both halves accumulate into the same fragment, q and k are split the same way
//      for the energy product (hh + hl + lh), so the 40x40 energies keep fp32-level accuracy: they sit in an
//      exponent. X, W_v, V and the attention weights are fp16 with fp32 accumulation.
//   2. warps (m, h), m = 16-row block of pixels, h = half of the channels: energy row block -> softmax in
//      registers -> the accumulator fragments are the A fragments of P V -> 1/rowsum -> staged fp32.
//   3. gamma * out + x, coalesced fp16 stores.
// Replaces the value-projection GEMM launch + the fp32 CUDA-core kernel (14 + 45 us per 640 frames).
struct PamMmaSmem {
  __half x[48][136];     // X[p][c], rows 40..47 zero
  __half v[48][136];     // V[p][c], rows 40..47 zero
  __half wqh[32][136];   // [W_q ; W_k] high halves
  __half wql[32][136];   // low halves
  __half wv[128][136];
  __half qh[48][40];     // q (cols 0..15) | k (cols 16..31), high halves
  __half ql[48][40];
  float bqk[32];
  float bv[128];
  float o[40][132];
};

__global__ void __launch_bounds__(256) pam_mma_kernel(const enc_t* __restrict__ xin, enc_t* __restrict__ out,
                                                      const float* __restrict__ wqk, const float* __restrict__ bqk,
                                                      const enc_t* __restrict__ wv, const float* __restrict__ bv,
                                                      float gamma, int B, int ldin) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t pam_raw2[];
  PamMmaSmem& s = *reinterpret_cast<PamMmaSmem*>(pam_raw2);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  // ---- once per CTA: weights (fp32 -> hi/lo fp16), biases, zero padding rows
  for (int i = tid; i < 32 * PAM_C; i += 256) {
    const int r = i >> 7, c = i & 127;
    const float w = wqk[i];
    const __half hi = __float2half_rn(w);
    s.wqh[r][c] = hi;
    s.wql[r][c] = __float2half_rn(w - __half2float(hi));
  }
  for (int i = tid; i < PAM_C * PAM_C / 8; i += 256) {
    const int r = i >> 4, c8 = (i & 15) * 8;
    *reinterpret_cast<uint4*>(&s.wv[r][c8]) = __ldg(reinterpret_cast<const uint4*>(wv + r * PAM_C + c8));
  }
  if (tid < 32) s.bqk[tid] = bqk[tid];
  if (tid < 128) {
    s.bv[tid] = bv[tid];
    *reinterpret_cast<uint4*>(&s.x[40 + (tid >> 4)][(tid & 15) * 8]) = make_uint4(0, 0, 0, 0);
    *reinterpret_cast<uint4*>(&s.v[40 + (tid >> 4)][(tid & 15) * 8]) = make_uint4(0, 0, 0, 0);
  }
  for (int f = blockIdx.x; f < B; f += gridDim.x) {
    __syncthreads();
    const enc_t* xf = xin + static_cast<long long>(f) * PAM_P * ldin;
    for (int i = tid; i < PAM_P * PAM_C / 8; i += 256) {
      const int p = i >> 4, c8 = (i & 15) * 8;
      *reinterpret_cast<uint4*>(&s.x[p][c8]) = __ldg(reinterpret_cast<const uint4*>(xf + p * ldin + c8));
    }
    __syncthreads();
    {  // ---- 1. projections: every warp two 8-column tiles of V, warps 0..3 also one tile of [q | k]
      float av[3][2][4], aq[3][4];
#pragma unroll
      for (int m = 0; m < 3; ++m) {
#pragma unroll
        for (int i = 0; i < 4; ++i) av[m][0][i] = av[m][1][i] = aq[m][i] = 0.f;
      }
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        uint32_t bvf[4], bqf[4];
        ldsm_x4(bvf, &s.wv[16 * warp + (lane & 7) + ((lane >> 4) & 1) * 8][16 * kk + ((lane >> 3) & 1) * 8]);
        if (warp < 4) {
          const int sel = lane >> 3;   // matrices: hi / k lo, hi / k hi, lo / k lo, lo / k hi
          const __half* wrow = (sel & 2) ? &s.wql[8 * warp + (lane & 7)][0] : &s.wqh[8 * warp + (lane & 7)][0];
          ldsm_x4(bqf, wrow + 16 * kk + (sel & 1) * 8);
        }
#pragma unroll
        for (int m = 0; m < 3; ++m) {
          uint32_t af[4];
          ldsm_x4(af, &s.x[16 * m + (lane & 7) + ((lane >> 3) & 1) * 8][16 * kk + ((lane >> 4) & 1) * 8]);
          mma_16816(av[m][0], af, bvf[0], bvf[1]);
          mma_16816(av[m][1], af, bvf[2], bvf[3]);
          if (warp < 4) {
            mma_16816(aq[m], af, bqf[0], bqf[1]);
            mma_16816(aq[m], af, bqf[2], bqf[3]);
          }
        }
      }
#pragma unroll
      for (int m = 0; m < 3; ++m) {
#pragma unroll
        for (int n = 0; n < 2; ++n) {
          const int c = 16 * warp + 8 * n + 2 * t;
          const float b0 = s.bv[c], b1 = s.bv[c + 1];
          const int r0 = 16 * m + g, r1 = r0 + 8;
          if (r0 < PAM_P) *reinterpret_cast<uint32_t*>(&s.v[r0][c]) = pack_h2(av[m][n][0] + b0, av[m][n][1] + b1);
          if (r1 < PAM_P) *reinterpret_cast<uint32_t*>(&s.v[r1][c]) = pack_h2(av[m][n][2] + b0, av[m][n][3] + b1);
        }
        if (warp < 4) {
          const int c = 8 * warp + 2 * t;
          const float b0 = s.bqk[c], b1 = s.bqk[c + 1];
#pragma unroll
          for (int hrow = 0; hrow < 2; ++hrow) {
            const int r = 16 * m + g + 8 * hrow;
            const float v0 = aq[m][2 * hrow] + b0, v1 = aq[m][2 * hrow + 1] + b1;
            const __half h0 = __float2half_rn(v0), h1 = __float2half_rn(v1);
            *reinterpret_cast<__half2*>(&s.qh[r][c]) = __halves2half2(h0, h1);
            *reinterpret_cast<__half2*>(&s.ql[r][c]) =
                __halves2half2(__float2half_rn(v0 - __half2float(h0)), __float2half_rn(v1 - __half2float(h1)));
          }
        }
      }
    }
    __syncthreads();
    if (warp < 6) {  // ---- 2. attention for pixel rows [16m, 16m+16), channels [64h, 64h+64)
      const int m = warp % 3, h = warp / 3;
      float e[6][4];
#pragma unroll
      for (int j = 0; j < 6; ++j) e[j][0] = e[j][1] = e[j][2] = e[j][3] = 0.f;
      uint32_t aqh[4], aql[4];
      ldsm_x4(aqh, &s.qh[16 * m + (lane & 7) + ((lane >> 3) & 1) * 8][((lane >> 4) & 1) * 8]);
      ldsm_x4(aql, &s.ql[16 * m + (lane & 7) + ((lane >> 3) & 1) * 8][((lane >> 4) & 1) * 8]);
#pragma unroll
      for (int jp = 0; jp < 3; ++jp) {
        uint32_t bkh[4], bkl[4];
        ldsm_x4(bkh, &s.qh[16 * jp + (lane & 7) + ((lane >> 4) & 1) * 8][16 + ((lane >> 3) & 1) * 8]);
        ldsm_x4(bkl, &s.ql[16 * jp + (lane & 7) + ((lane >> 4) & 1) * 8][16 + ((lane >> 3) & 1) * 8]);
        mma_16816(e[2 * jp], aqh, bkh[0], bkh[1]);
        mma_16816(e[2 * jp], aqh, bkl[0], bkl[1]);
        mma_16816(e[2 * jp], aql, bkh[0], bkh[1]);
        mma_16816(e[2 * jp + 1], aqh, bkh[2], bkh[3]);
        mma_16816(e[2 * jp + 1], aqh, bkl[2], bkl[3]);
        mma_16816(e[2 * jp + 1], aql, bkh[2], bkh[3]);
      }
      e[5][0] = e[5][1] = e[5][2] = e[5][3] = -INFINITY;   // key positions 40..47 do not exist
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        mx[0] = fmaxf(mx[0], fmaxf(e[j][0], e[j][1]));
        mx[1] = fmaxf(mx[1], fmaxf(e[j][2], e[j][3]));
      }
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        mx[0] = fmaxf(mx[0], __shfl_xor_sync(0xffffffffu, mx[0], o));
        mx[1] = fmaxf(mx[1], __shfl_xor_sync(0xffffffffu, mx[1], o));
      }
      float sum[2] = {0.f, 0.f};
#pragma unroll
      for (int j = 0; j < 6; ++j) {
        e[j][0] = __expf(e[j][0] - mx[0]), e[j][1] = __expf(e[j][1] - mx[0]);
        e[j][2] = __expf(e[j][2] - mx[1]), e[j][3] = __expf(e[j][3] - mx[1]);
        sum[0] += e[j][0] + e[j][1], sum[1] += e[j][2] + e[j][3];
      }
#pragma unroll
      for (int o = 1; o <= 2; o <<= 1) {
        sum[0] += __shfl_xor_sync(0xffffffffu, sum[0], o);
        sum[1] += __shfl_xor_sync(0xffffffffu, sum[1], o);
      }
      float oacc[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j) oacc[j][0] = oacc[j][1] = oacc[j][2] = oacc[j][3] = 0.f;
#pragma unroll
      for (int kk = 0; kk < 3; ++kk) {
        uint32_t af[4];
        af[0] = pack_h2(e[2 * kk][0], e[2 * kk][1]), af[1] = pack_h2(e[2 * kk][2], e[2 * kk][3]);
        af[2] = pack_h2(e[2 * kk + 1][0], e[2 * kk + 1][1]), af[3] = pack_h2(e[2 * kk + 1][2], e[2 * kk + 1][3]);
#pragma unroll
        for (int cp = 0; cp < 4; ++cp) {
          uint32_t bf[4];
          ldsm_x4_trans(bf, &s.v[16 * kk + (lane & 7) + ((lane >> 3) & 1) * 8][64 * h + 16 * cp + ((lane >> 4) & 1) * 8]);
          mma_16816(oacc[2 * cp], af, bf[0], bf[1]);
          mma_16816(oacc[2 * cp + 1], af, bf[2], bf[3]);
        }
      }
      const float inv[2] = {1.f / sum[0], 1.f / sum[1]};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int c = 64 * h + 8 * j + 2 * t;
        const int r0 = 16 * m + g, r1 = r0 + 8;
        if (r0 < PAM_P) *reinterpret_cast<float2*>(&s.o[r0][c]) = make_float2(oacc[j][0] * inv[0], oacc[j][1] * inv[0]);
        if (r1 < PAM_P) *reinterpret_cast<float2*>(&s.o[r1][c]) = make_float2(oacc[j][2] * inv[1], oacc[j][3] * inv[1]);
      }
    }
    __syncthreads();
    enc_t* of = out + static_cast<long long>(f) * PAM_P * PAM_C;
    for (int i = tid; i < PAM_P * PAM_C / 8; i += 256) {
      const int p = i >> 4, c8 = (i & 15) * 8;
      const float4 o_lo = *reinterpret_cast<const float4*>(&s.o[p][c8]);
      const float4 o_hi = *reinterpret_cast<const float4*>(&s.o[p][c8 + 4]);
      const uint4 ux = *reinterpret_cast<const uint4*>(&s.x[p][c8]);
      const __half* hx = reinterpret_cast<const __half*>(&ux);
      uint4 u;
      u.x = enc_pack2(gamma * o_lo.x + __half2float(hx[0]), gamma * o_lo.y + __half2float(hx[1]));
      u.y = enc_pack2(gamma * o_lo.z + __half2float(hx[2]), gamma * o_lo.w + __half2float(hx[3]));
      u.z = enc_pack2(gamma * o_hi.x + __half2float(hx[4]), gamma * o_hi.y + __half2float(hx[5]));
      u.w = enc_pack2(gamma * o_hi.z + __half2float(hx[6]), gamma * o_hi.w + __half2float(hx[7]));
      *reinterpret_cast<uint4*>(of + p * PAM_C + c8) = u;
    }
  }
}

void launch_pam_mma(const enc_t* x, enc_t* out, const float* wqk, const float* bqk, const enc_t* wv, const float* bv,
                    float gamma, int B, int ldin, int num_sms, cudaStream_t stream) {
  static size_t cfg[CADRE_MAX_DEVICES] = {};
  ensure_dynamic_smem(pam_mma_kernel, sizeof(PamMmaSmem), cfg);
  const int grid = B < 2 * num_sms ? B : 2 * num_sms;
  launch_k(pam_mma_kernel, dim3(grid), dim3(256), sizeof(PamMmaSmem), stream, x, out, wqk, bqk, wv, bv, gamma, B, ldin);
  CADRE_CUDA_CHECK(cudaGetLastError());
}
#endif

void launch_cam(const enc_t* x, enc_t* out, float gamma, int B, int ldin, int num_sms,
                cudaStream_t stream) {
  static size_t cfg[CADRE_MAX_DEVICES] = {};
  ensure_dynamic_smem(cam_mma_kernel, sizeof(CamMmaSmem), cfg);
  const int grid = B < 6 * num_sms ? B : 6 * num_sms;
  launch_k(cam_mma_kernel, dim3(grid), dim3(256), sizeof(CamMmaSmem), stream, x, out, gamma, B, ldin);
  CADRE_CUDA_CHECK(cudaGetLastError());
}

// ---------------------------------------------------------------------------------------------------------
// InterTaskAtt.forward, transformer branch (intertask_att.py:137-176): rank-1 energies
//   att_bc[i]  = sum_j v_bc[j]  softmax_j((q_vis[i]/16) k_bc[j])  + v_bc[i]
//   att_vis[i] = sum_j v_vis[j] softmax_j((q_bc[i]/16)  k_vis[j]) + v_vis[i]
// qkv: fp32 [6][B][256] in the order (vis q, vis k, vis v, bc q, bc k, bc v). Output row: [att_vis | att_bc |
// optional 18 measurement floats] fp32 with row stride ld_out (danet.py:233 + agent.py:106-111).
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(256) intertask_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                        const double* __restrict__ meas, int B, int ld_out) {
  pdl_trigger();
  pdl_wait();
  __shared__ __align__(16) float sk[2][256], sv[2][256];
  __shared__ float s_ext[2][2][8];   // [direction][max, min][warp]
  const int f = blockIdx.x, i = threadIdx.x;
  const long long BS = static_cast<long long>(B) * 256;
  const float* base = qkv + static_cast<long long>(f) * 256;
  const float vq = base[0 * BS + i], vk = base[1 * BS + i], vv = base[2 * BS + i];
  const float bq = base[3 * BS + i], bk = base[4 * BS + i], bv = base[5 * BS + i];
  sk[0][i] = bk, sv[0][i] = bv;  // direction 0: visual query -> bc keys/values  => att_bc
  sk[1][i] = vk, sv[1][i] = vv;  // direction 1: bc query -> visual keys/values  => att_vis
  {  // extrema of the keys: max_j(q k_j) = q * (q >= 0 ? kmax : kmin), so no per-thread max pass is needed
    float mx0 = bk, mn0 = bk, mx1 = vk, mn1 = vk;
    for (int o = 16; o > 0; o >>= 1) {
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, o)), mn0 = fminf(mn0, __shfl_xor_sync(0xffffffffu, mn0, o));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, o)), mn1 = fminf(mn1, __shfl_xor_sync(0xffffffffu, mn1, o));
    }
    if ((i & 31) == 0) {
      s_ext[0][0][i >> 5] = mx0, s_ext[0][1][i >> 5] = mn0;
      s_ext[1][0][i >> 5] = mx1, s_ext[1][1][i >> 5] = mn1;
    }
  }
  __syncthreads();
  constexpr float LOG2E = 1.4426950408889634f;
  const float qs[2] = {vq / 16.0f, bq / 16.0f};
  float res[2];
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    const float q = qs[d];
    float kmax = s_ext[d][0][0], kmin = s_ext[d][1][0];
#pragma unroll
    for (int w = 1; w < 8; ++w) kmax = fmaxf(kmax, s_ext[d][0][w]), kmin = fminf(kmin, s_ext[d][1][w]);
    const float ql = q * LOG2E;                         // softmax in base 2: exp(x - m) = 2^(x*log2e - m*log2e)
    const float ml = ql * (q >= 0.f ? kmax : kmin);
    float sum = 0.f, acc = 0.f;
#pragma unroll 4
    for (int j = 0; j < 256; j += 4) {
      const float4 k4 = *reinterpret_cast<const float4*>(&sk[d][j]);
      const float4 v4 = *reinterpret_cast<const float4*>(&sv[d][j]);
      const float p0 = ex2_approx(fmaf(ql, k4.x, -ml)), p1 = ex2_approx(fmaf(ql, k4.y, -ml));
      const float p2 = ex2_approx(fmaf(ql, k4.z, -ml)), p3 = ex2_approx(fmaf(ql, k4.w, -ml));
      sum += (p0 + p1) + (p2 + p3);
      acc = fmaf(p0, v4.x, fmaf(p1, v4.y, fmaf(p2, v4.z, fmaf(p3, v4.w, acc))));
    }
    res[d] = acc / sum + sv[d][i];
  }
  float* o = out + static_cast<long long>(f) * ld_out;
  o[i] = res[1];         // att_visual first (danet.py:233 cat((att_visual, att_bc)))
  o[256 + i] = res[0];
  if (meas != nullptr && i < 18) o[512 + i] = static_cast<float>(meas[static_cast<long long>(f) * 3 + (i % 3)]);
}

void launch_intertask(const float* qkv, float* out, const double* meas, int B, int ld_out,
                      cudaStream_t stream) {
  launch_k(intertask_kernel, dim3(B), dim3(256), 0, stream, qkv, out, meas, B, ld_out);
  CADRE_CUDA_CHECK(cudaGetLastError());
}

}  // namespace cadre

/*
 * cadre_b200.h — C ABI of libcadre_sm100.so, the sm_100a device side of the CADRE learner hot path.
 *
 * The reference (BIT-MCS/Cadre) is pure Python/PyTorch and has no FFI of its own; each entry
 * point below names the reference call site it replaces. Conventions:
 *   - every function returns 0 on success, non-zero on error; cadre_last_error() returns a thread-local
 *     message; no C++ exception crosses the boundary;
 *   - pointer arguments are raw DEVICE pointers borrowed from the caller (e.g. torch.Tensor.data_ptr()) unless
 *     the name ends in _host; the library never frees them; the caller keeps them alive until the stream
 *     has drained;
 *   - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream().cuda_stream);
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails with an error.
 */
#ifndef CADRE_B200_H_
#define CADRE_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* cadre_last_error(void);
/* library / build identification: "cadre_b200 sm_100a <date>" */
const char* cadre_version(void);

/* ------------------------------------------------------------------------------------------------------
 * Dense tile operator (tcgen05 + TMEM + TMA). Replaces the cuBLAS/cuDNN calls PyTorch dispatches for
 * nn.Linear / nn.LSTMCell / F.conv2d on the path (ppo_agent/models.py:139-152,165-212;
 * carla_perception/Networks/danet_blocks/resnet.py:39-55,168-183; intertask_att.py:39-80).
 * out[M,N] = epilogue(alpha * A * B^T); kind 0: bf16 operands, kind 1: fp32 operands consumed as TF32.
 * a_mn / b_mn = 0: operand stored [M|N][K] (K contiguous); 1: stored [K][M|N] (M|N contiguous).
 * All leading dimensions / batch strides are in elements. */
typedef struct cadre_gemm_args {
  int32_t kind, a_mn, b_mn, batch;
  int32_t M, N, K, block_n;
  const void* A;
  const void* B;
  int64_t lda, a_bs, ldb, b_bs;
  void* out;
  int64_t ldc, out_bs;
  int32_t out_f32, act; /* act: 0 none, 1 ReLU, 2 LeakyReLU(0.01) */
  const float* bias;
  int64_t bias_bs;
  const void* res; /* residual, dtype of out */
  int64_t ldr, res_bs;
  const void* mask; /* out = 0 where mask <= 0, dtype of out */
  int64_t ldm, mask_bs;
  int32_t res_after_act, rows_is_k;
  const int32_t* batch_rows; /* optional per-batch valid rows (or valid K if rows_is_k) */
  float alpha;
  int32_t epi; /* 0 linear, 1 LSTM cell */
  const float* xpart;
  const float* c_prev;
  float* c_out;
  float* h_out;
  float* gates_out;
  int64_t ldx, x_bs, ldh, h_bs;
} cadre_gemm_args;
int cadre_gemm(const cadre_gemm_args* args, void* stream);

/* NHWC bf16 implicit-GEMM convolution with folded BatchNorm bias, optional residual and ReLU
 * (resnet.py:39-55 BasicBlock, danet.py:21-36 conv5a/5c/51/52, danet.py:41 conv8). */
int cadre_conv2d_nhwc(const void* in, int B, int Hin, int Win, int Cin, const void* w, int Cout, int KH,
                      int KW, int stride, int pad, const float* bias, const void* res, int res_after_act,
                      int act, void* out, void* stream);

/* ResNet stem conv7x7/s2/p3 + folded BN + ReLU over the padded 4-channel image written by
 * cadre_preprocess (resnet.py:169-171). */
int cadre_stem_conv(const void* in_padded, int B, const void* w256, const float* bias, void* out,
                    void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CADRE_B200_H_ */

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
/* 16-bit storage / tensor-core operand type of the encoder path: 1 = IEEE fp16 (default build), 0 = bf16.
 * Every "enc16" buffer below (activations, conv / linear weights) uses this type. */
int cadre_enc_dtype(void);
/* stream-ordered device-to-device copy (test / debug helper) */
int cadre_memcpy_d2d(void* dst, const void* src, int64_t nbytes, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Dense tile operator (tcgen05 + TMEM + TMA). Replaces the cuBLAS/cuDNN calls PyTorch dispatches for
 * nn.Linear / nn.LSTMCell / F.conv2d on the path (ppo_agent/models.py:139-152,165-212;
 * carla_perception/Networks/danet_blocks/resnet.py:39-55,168-183; intertask_att.py:39-80).
 * out[M,N] = epilogue(alpha * A * B^T); kind 0: enc16 operands, kind 1: fp32 operands consumed as TF32.
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

/* NHWC enc16 implicit-GEMM convolution with folded BatchNorm bias, optional residual and ReLU
 * (resnet.py:39-55 BasicBlock, danet.py:21-36 conv5a/5c/51/52, danet.py:41 conv8). in_pad = 1: the input tensor is
 * stored with a 1-pixel zero border, [B][Hin+2][Win+2][Cin] (Hin / Win stay the logical sizes). */
int cadre_conv2d_nhwc(const void* in, int B, int Hin, int Win, int Cin, const void* w, int Cout, int KH,
                      int KW, int stride, int pad, const float* bias, const void* res, int res_after_act,
                      int act, void* out, int in_pad, void* stream);

/* Halo-reuse 3x3 / stride 1 / pad 1 convolution, 64 -> 64 channels (ResNet layer1), on zero-bordered
 * activations: in / res / out are [B][H+2][W+2][64] enc16 with zero borders (kept zero by the kernel). */
/* Debug: per-CTA cycle counters ([grid][16] int64, device memory) filled by the layer1 / stem kernels while
 * the pointer is set; pass NULL to switch off. Not part of the reference interface. */
int cadre_debug_clk(long long* dev_counters);

int cadre_conv3x3_flat64(const void* in, int B, int H, int W, const void* w, const float* bias, const void* res,
                         int act, void* out, void* stream);
/* The same for 128 -> 128 channels (ResNet layer2, W <= 33): in / res / out are [B][H+2][W+2][128] enc16 with zero
 * borders, w is [128][3][3][128]; res (optional) is added before the activation (resnet.py:52-53). */
int cadre_conv3x3_flat128(const void* in, int B, int H, int W, const void* w, const float* bias, const void* res,
                          int act, void* out, void* stream);
/* cadre_conv2d_nhwc writing into a zero-bordered output tensor [B][Hout+2][Wout+2][Cout] (the border is left
 * untouched: the caller zeroes it once). Feeds the halo-reuse kernels above. */
int cadre_conv2d_nhwc_bordered_out(const void* in, int B, int Hin, int Win, int Cin, const void* w, int Cout, int KH,
                                   int KW, int stride, int pad, const float* bias, int act, void* out, int in_pad,
                                   void* stream);

/* ------------------------------------------------------------------------------------------------------
 * Perception encoder forward = DANet.get_latent_feature(x, "concate")
 * (carla_perception/Networks/danet.py:216-238) + CadreAgent.pre_process / get_latent_feature
 * (ppo_agent/agent.py:43-75, 97-112). Weights are prepared once by the host (BatchNorm folded into the
 * preceding conv, NHWC / K-major bf16 layouts, linear chains folded) — see cadre_b200/encoder.py. */
typedef struct cadre_encoder_weights {
  const void* stem_w;   /* enc16 [64][256]: [cout][row pair 4][kw 8][row 2][c 4], zero where kh==7 or kw==7 */
  const float* stem_b;  /* [64] conv bias and BN folded */
  const void* conv_w[19]; /* backbone convs in execution order (per block: conv1, conv2, [downsample]);
                             enc16 [Cout][KH][KW][Cin], BN folded */
  const float* conv_b[19];
  const void* head5_w;  /* conv5a | conv5c stacked on Cout: enc16 [256][3][3][512] */
  const float* head5_b; /* [256] */
  const float* pam_wqk; /* fp32 [32][128]: query_conv rows 0..15, key_conv rows 16..31 */
  const float* pam_bqk; /* [32] */
  const void* pam_wv;   /* enc16 [128][128] value_conv (runs on the tensor cores) */
  const float* pam_bv;  /* [128] */
  const void* conv51_w; /* enc16 [128][3][3][128] */
  const float* conv51_b;
  const void* conv52_w;
  const float* conv52_b;
  const void* fc1_w;    /* enc16 [3072][5120]: Linear(20480,512) x6 o {visual,bc}_conv o conv8, K = (h*8+w)*128+c */
  const float* fc1_b;   /* [3072] */
  const void* fc2_w;    /* enc16 [6][256][512], order vis q,k,v, bc q,k,v */
  const float* fc2_b;   /* [6][256] */
  float pam_gamma, cam_gamma;
} cadre_encoder_weights;

int cadre_encoder_create(void** handle, const cadre_encoder_weights* w, int max_batch);
int cadre_encoder_destroy(void* handle);
/* rgb u8 [B][144][256][3], route_fig u8 [B][256][144], measurements f64 [B][3] or NULL;
 * out fp32 [B][ld_out]: 512 latent floats (+ 18 measurement floats when measurements != NULL). */
int cadre_encoder_forward_u8(void* handle, const uint8_t* rgb, const uint8_t* route_fig,
                             const double* measurements, int B, float* out, int ld_out, void* stream);
/* x fp32 NCHW [B][4][144][256] already pre-processed (BASELINE config 2, encoder sweep). */
int cadre_encoder_forward_f32(void* handle, const float* x_nchw, int B, float* out, int ld_out, void* stream);
/* internal activation buffers for parity tests: 0 layer4 out, 1 feat_sum, 2 conv5a|5c, 3 PAM, 4 CAM, 5 stem */
int cadre_encoder_buffer(void* handle, int which, void** ptr, int64_t* elems_per_frame);
/* Re-runs the trunk on the frames ingested by the previous forward call with a CUDA event between launches:
 * ms_out[i] = device time of launch i (<= 64), names_out = ';'-joined launch names. Measurement helper for
 * bench.py's roofline object; synchronises the stream. */
int cadre_encoder_profile(void* handle, int B, float* out, int ld_out, float* ms_out, char* names_out,
                          int names_cap, int* n_out, void* stream);
/* kernels launched by the last forward call */
int cadre_encoder_launches(void* handle);

/* ------------------------------------------------------------------------------------------------------
 * Rollout: RolloutStorage.compute_returns GAE branch (ppo_agent/storage.py:68-76) + advantage
 * normalisation with unbiased std (ppo_agent/train.py:82-88), one sequence per (env, head).
 * rewards / values / masks / returns: fp32 [E][T+1]; next_value fp32 [E]; adv fp32 [E][T].
 * values[e][T] is overwritten with next_value[e] exactly as the reference does. */
int cadre_gae(const float* rewards, float* values, const float* masks, const float* next_value, float* returns,
              float* adv, int E, int T, float gamma, float tau, int normalize, void* stream);

/* Sliding-window assembly of rollout observations (env_wrapper.py:900-914 stacks the last `seq_length` frames per
 * tick; train.py:69-72 inserts that window into the steer AND the throttle storage). unique_feats: fp32
 * [workers][num_steps + seq_length - 1][feature_dims] = encoder features of every DISTINCT frame of each worker's
 * rollout, oldest first. Writes obs[(w*2 + head)*obs_head_stride + t*obs_step_stride + j*feature_dims + :] =
 * unique_feats[w][t + j][:] for head 0..1, t < num_steps, j < seq_length (strides in floats; RolloutStorage.obs of
 * the batched pool: head stride (T+1)*seq*F, step stride seq*F). */
int cadre_window_scatter(const float* unique_feats, float* obs, int workers, int num_steps, int seq_length,
                         int feature_dims, int64_t obs_head_stride, int64_t obs_step_stride, void* stream);

/* ------------------------------------------------------------------------------------------------------
 * PPO update = CadreAgent.update_policy (ppo_agent/agent.py:166-237) for `workers` logical workers at once
 * (forward + hand-written backward; gradients are the SUM over workers, like Shared_grad_buffers.add_gradient,
 * ppo_agent/models.py:231-239), then the chief's per-module clip + Adam (ppo_agent/chief.py:13-21).
 * Parameters / gradients / Adam moments are flat fp32 buffers of cadre_ppo_param_count() elements in the
 * layout of cadre_b200/csrc/ppo_layout.h (converted from / to reference state dicts by
 * cadre_b200/ppo_params.py). Between cadre_ppo_update and cadre_ppo_adam_step the host all-reduces (sum) the
 * gradient buffer across ranks (NCCL), which replaces the shared-memory aggregation of the reference. */
typedef struct cadre_ppo_config {
  int32_t workers;    /* logical workers (envs) whose minibatches are processed per update on this GPU */
  int32_t mini_batch; /* rows per worker per head = num_steps / mini_batch_num (storage.py:94) */
  float clip, value_coeff, clip_coeff, ent_coeff; /* agent_config.py:43-46 */
} cadre_ppo_config;

/* device pointers into one RolloutStorage (ppo_agent/storage.py:8-26) + its advantage vector */
typedef struct cadre_storage_ref {
  const float* obs;              /* [T+1][8][530] */
  const int64_t* action;         /* [T+1][1] */
  const float* value_preds;      /* [T+1][1] */
  const float* returns;          /* [T+1][1] */
  const float* action_log_probs; /* [T+1][1] */
  const float* adv;              /* [T][1] normalised advantages */
  const float* hn;               /* [T+1][530] */
  const float* cn;               /* [T+1][530] */
  const int32_t* command;        /* [T+1][1] */
} cadre_storage_ref;

int64_t cadre_ppo_param_count(void);
int cadre_ppo_create(void** handle, const cadre_ppo_config* cfg);
int cadre_ppo_destroy(void* handle);
/* storages_host: [workers][2] (steer, throttle); indices_host: int32 [workers][2][mini_batch] minibatch row
 * indices (host-generated, bit-exact torch.randperm chunks); losses: device fp32 [workers][2][3] =
 * per-worker, per-head UN-scaled (value, action, entropy) losses. grads is overwritten.
 * Staged form for a sequence of update steps on the same storages (train.py:93-110: ppo_epoch x minibatches):
 * cadre_ppo_stage uploads the storage references and the indices of n_steps <= 64 steps ([n_steps][workers][2][mb])
 * once; each following cadre_ppo_update(handle, NULL, NULL, ...) consumes the next slice (a device-side counter) and
 * advances the device-side Adam step (starting at first_adam_step), which cadre_ppo_adam_step* use when called with
 * step = 0. Such an update + adam sequence issues no host-to-device copy and can be captured in a CUDA graph. */
int cadre_ppo_stage(void* handle, const cadre_storage_ref* storages_host, const int32_t* indices_host, int n_steps,
                    int first_adam_step, void* stream);
int cadre_ppo_update(void* handle, const cadre_storage_ref* storages_host, const int32_t* indices_host,
                     float* params, float* grads, float* losses, void* stream);
/* Forward only (CadreAgent.act / get_value, Model.evaluate_actions; agent.py:114-164, models.py:184-212):
 * row_out fp32 [2 heads][workers*mini_batch][36] = value, log-prob(stored action), entropy, 33 normalised
 * logits (first 3 valid for the throttle head); row order = (worker, minibatch position). */
int cadre_ppo_evaluate(void* handle, const cadre_storage_ref* storages_host, const int32_t* indices_host,
                       const float* params, float* row_out, void* stream);
int cadre_ppo_adam_step(void* handle, float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                        float max_grad_norm, float lr, float beta1, float beta2, float eps, int step,
                        void* stream);
/* gradient norms of the 16 modules seen by the last adam_step: [0..7] LSTM of expert e, [8..15] actor-critic
 * of expert e (expert = head*4 + command); synchronises */
int cadre_ppo_module_norms(void* handle, float* norms16_host);
/* Data-parallel pipeline (replaces Shared_grad_buffers.add_gradient + the chief's single optimizer.step(),
 * models.py:231-239, chief.py:13-21, by per-group all-reduce + clip + Adam overlapped with the backward pass).
 * cadre_ppo_set_grad_groups(groups in {1, 2, 4, 8}): subsequent cadre_ppo_update calls produce the LSTM weight
 * gradients per group of 8 / groups experts. cadre_ppo_grad_range: the contiguous [offset, offset + count) range of the
 * flat buffers that holds group `group`'s LSTM tensors (group = -1: the actor-critic tensors of all experts) and the
 * module ids [mod_begin, mod_end) it covers (module e < 8 = LSTM of expert e, 8 + e = actor-critic of expert e: each
 * is one reference nn.Module, the unit of clip_grad_norm_). cadre_ppo_wait_grads: `stream` waits until the LAST
 * cadre_ppo_update has finished that range of `grads`. cadre_ppo_adam_step_modules: clip + Adam restricted to the
 * modules [mod_begin, mod_end). */
int cadre_ppo_set_grad_groups(void* handle, int groups);
int cadre_ppo_grad_range(void* handle, int group, int64_t* offset, int64_t* count, int* mod_begin, int* mod_end);
int cadre_ppo_wait_grads(void* handle, int group, void* stream);
int cadre_ppo_adam_step_modules(void* handle, float* params, const float* grads, float* exp_avg, float* exp_avg_sq,
                                float max_grad_norm, float lr, float beta1, float beta2, float eps, int step,
                                int mod_begin, int mod_end, void* stream);
/* Synchronises and returns non-zero (cadre_last_error() explains) if a persistent LSTM recurrence kernel of an
 * earlier cadre_ppo_update / cadre_ppo_evaluate on this handle gave up waiting for a cross-CTA hand-off (its polling is
 * bounded so that a lost signal can never hang the GPU). */
int cadre_ppo_check(void* handle);
int cadre_ppo_launches(void* handle);

/* ------------------------------------------------------------------------------------------------------
 * Gradient exchange across the ranks of one NVSwitch domain, in place: grads <- sum over ranks, every rank receives
 * the same bits. Replaces Shared_grad_buffers.add_gradient + the chief installing the summed gradient
 * (ppo_agent/models.py:231-239, ppo_agent/chief.py:13-16). The buffer must be symmetric memory: buffer_ptrs_host[r] =
 * rank r's copy as mapped into THIS process (r == rank: the local buffer), `multicast` = the multicast mapping of all
 * copies (NULL when the fabric has none), flag_ptrs_host[r] = rank r's zero-initialised flag area of
 * cadre_allreduce_flag_bytes() bytes, also symmetric memory (cadre_b200/collective.py obtains all of them from
 * torch.distributed._symmetric_memory). cadre_allreduce_sum reduces [offset, offset + count) (floats, multiples of 4);
 * all ranks must issue the same sequence of calls. use_multicast = 1: in-switch reduction (multimem.ld_reduce /
 * multimem.st); 0: peer loads summed in rank order + peer stores. cadre_allreduce_check synchronises and fails if a
 * barrier ever timed out (polling is bounded; a lost peer cannot hang the GPU). */
int cadre_allreduce_flag_bytes(void);
int cadre_allreduce_create(void** handle, int rank, int world, float* const* buffer_ptrs_host, float* multicast,
                           uint32_t* const* flag_ptrs_host, int64_t count);
/* thread blocks of 256 threads per launch (default 32, at most 256): every rank must use the same value for the same call */
int cadre_allreduce_set_blocks(void* handle, int blocks);
int cadre_allreduce_destroy(void* handle);
int cadre_allreduce_sum(void* handle, int64_t offset, int64_t count, int use_multicast, void* stream);
int cadre_allreduce_check(void* handle);

#ifdef __cplusplus
}
#endif
#endif /* CADRE_B200_H_ */

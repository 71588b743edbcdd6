/*
 * ivosw_b200 — C ABI of the B200-native frame-scoring path of IVOS-W.
 *
 * The reference (svip-lab/IVOS-W) is pure Python and has no FFI of its own; the
 * boundary it exposes for this path is three Python call signatures
 * (SURVEY.md §8(b)):
 *   B1  models/agent.py:168        Agent.action(state[T,2]) -> int
 *       models/agent.py:33         Brain.forward(N x T x 2) -> N x T
 *   B2  models/assessment.py:164   AssessNet.forward(tf B x 3 x H x W, tp B x H x W) -> B x 1
 *   a1  utils/utils_agent.py:77    recommend_frame(...)  (setting='wild', method='ours', lines 111-122)
 *   a10 utils/utils_manet.py:76-81,160-161  upsample / argmax / softmax tail of get_results
 * Each entry point below names the reference interface it sits under.  The
 * Python drop-in modules (ivos-w_b200/dropin/{models,utils}/) keep those
 * signatures and call this library through ctypes; INTEGRATION.md shows the
 * binding.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / C++ types cross the boundary;
 *   - *_dev pointers are caller-owned device memory on the context's device,
 *     *_host pointers are caller-owned host memory (pinned or pageable);
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*, NULL = the
 *     legacy default stream) and is NOT synchronised unless stated;
 *   - the only memory the library owns is the context's weights + workspace,
 *     released by ivosw_destroy();
 *   - every function returns IVOSW_OK (0) or an error code; the message is
 *     available (thread-local) from ivosw_last_error().  An allocation failure
 *     returns IVOSW_ERR_OOM and a message containing "out of memory", which
 *     the Python shim re-raises as RuntimeError so that the reference's
 *     OOM-retry loop (eval_agent_manet.py:382-396) keeps working.
 *   - there is no CPU fallback anywhere behind this ABI.
 */
#ifndef IVOSW_B200_H
#define IVOSW_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IVOSW_ABI_VERSION 1

#if defined(__GNUC__)
#define IVOSW_API __attribute__((visibility("default")))
#else
#define IVOSW_API
#endif

enum ivosw_status {
    IVOSW_OK = 0,
    IVOSW_ERR_INVALID = 1, /* bad argument / shape */
    IVOSW_ERR_CUDA = 2,    /* CUDA runtime / driver error */
    IVOSW_ERR_OOM = 3,     /* cudaErrorMemoryAllocation: message contains "out of memory" */
    IVOSW_ERR_STATE = 4    /* weights not loaded, wrong device, ... */
};

/* How the ResNet-50 convolutions of AssessNet are evaluated. */
enum ivosw_conv_mode {
    IVOSW_CONV_SIMT_FP32 = 0, /* fp32 CUDA-core implicit GEMM: validation path                     */
    IVOSW_CONV_TC_FP16X3 = 1, /* tcgen05 kind::f16, split-fp16 operands, 3 MMA terms: fp32-grade   */
    IVOSW_CONV_TC_FP16X1 = 2  /* tcgen05 kind::f16, single fp16 term: fast, ~1e-3 relative         */
};

typedef struct ivosw_ctx ivosw_ctx;

IVOSW_API int ivosw_abi_version(void);
IVOSW_API const char* ivosw_last_error(void);

/* Creates a context on CUDA device `device`. */
IVOSW_API int ivosw_create(int device, int conv_mode, ivosw_ctx** out);
IVOSW_API void ivosw_destroy(ivosw_ctx* ctx);
IVOSW_API int ivosw_set_conv_mode(ivosw_ctx* ctx, int conv_mode);
/* Number of kernels this context has launched since creation (bench.py's gpu_launches). */
IVOSW_API long long ivosw_launch_count(const ivosw_ctx* ctx);
/* Host-to-device bytes the last host-buffer call (ivosw_round_host / ivosw_score_shard_host) actually sent: the
 * background probability channel and the frame rows no ROI can touch are not transferred (bench.py's
 * e2e.h2d_bytes_per_step). */
IVOSW_API long long ivosw_last_h2d_bytes(const ivosw_ctx* ctx);

/* ---- Q-network (models/agent.py::Brain) --------------------------------------------------
 * params_host: the 180 993 floats of Brain.state_dict() concatenated in this order
 * (row-major, torch layout): encoder_fc1.weight[128x2], encoder_fc1.bias[128],
 * encoder_fc2.weight[128x128], encoder_fc2.bias[128], lstm_cell.weight_ih[512x128],
 * lstm_cell.weight_hh[512x128], decoder_fc1.weight[128x256], decoder_fc1.bias[128],
 * decoder_fc2.weight[1x128], decoder_fc2.bias[1]. */
#define IVOSW_BRAIN_NUM_PARAMS 180993
IVOSW_API int ivosw_brain_load(ivosw_ctx* ctx, const float* params_host, size_t n_floats);

/* Brain.forward (models/agent.py:33-64) + the first-max argmax of Agent.action (:187-188).
 * state_dev: N x T x 2 fp32; q_dev: N x T fp32; argmax_dev: N int32 (may be NULL). */
IVOSW_API int ivosw_brain_forward(ivosw_ctx* ctx, const float* state_dev, int N, int T,
                        float* q_dev, int* argmax_dev, void* stream);

/* ---- Double-DQN training step (models/agent.py::Agent.update_agent, :103-166; BASELINE config C4) ----
 * The policy network is the one loaded with ivosw_brain_load (updated IN PLACE); the target network is
 * loaded with ivosw_dqn_load_target or copied from the policy with ivosw_dqn_sync_target (the caller
 * draws np.random.random() < update_rate, agent.py:163-165).  ivosw_dqn_update runs: no-grad policy /
 * target forwards on new_state, forward + backward on state, element-wise gradient clamp to +-1, Adam
 * (betas 0.9 / 0.999, eps 1e-8, L2 weight decay in the gradient).  All *_dev arrays are device memory:
 * state / new_state N x T x 2 fp32, action N int32, rewards N fp32.  loss_host receives the scalar loss
 * (the call synchronises); grads_dev (nullable, 180 993 floats, blob order) the clamped gradients. */
IVOSW_API int ivosw_dqn_load_target(ivosw_ctx* ctx, const float* params_host, size_t n_floats);
IVOSW_API int ivosw_dqn_sync_target(ivosw_ctx* ctx, void* stream);
IVOSW_API int ivosw_dqn_reset_optimizer(ivosw_ctx* ctx);
/* Adam state for checkpoints (models/agent.py:101 optimizer; the reference's torch.optim.Adam.state_dict()):
 * exp_avg / exp_avg_sq as 180 993 floats in blob order (device memory) and the step count.  Loading new policy
 * weights from the host (ivosw_brain_load) resets this state; ivosw_dqn_set_optimizer restores it afterwards. */
IVOSW_API int ivosw_dqn_get_optimizer(ivosw_ctx* ctx, float* m_dev, float* v_dev, long long* step_out, void* stream);
IVOSW_API int ivosw_dqn_set_optimizer(ivosw_ctx* ctx, const float* m_dev, const float* v_dev, long long step,
                                      void* stream);
IVOSW_API int ivosw_dqn_update(ivosw_ctx* ctx, const float* state_dev, const float* new_state_dev,
                     const int* action_dev, const float* reward_step_dev, const float* reward_done_dev,
                     int N, int T, float gamma, float lr, float weight_decay, float* loss_host,
                     float* grads_dev, int apply_update, void* stream);
/* Data-parallel form: call ivosw_dqn_update(apply_update = 0) on every rank's slice of the batch (grads_dev
 * then receives the RAW gradient of that slice's mean loss), average grads_dev across ranks (one NCCL
 * all-reduce of 724 KB), then ivosw_dqn_apply clamps to +-1 and takes the Adam step — the same arithmetic
 * as one full-batch update.  grads_dev is overwritten with the clamped gradient. */
IVOSW_API int ivosw_dqn_apply(ivosw_ctx* ctx, float* grads_dev, float lr, float weight_decay, void* stream);
/* Copies a parameter set (0 = policy, 1 = target) to out_dev (180 993 floats, blob order). */
IVOSW_API int ivosw_brain_get_params(ivosw_ctx* ctx, int which, float* out_dev, void* stream);

/* ---- quality CNN (models/assessment.py::AssessNet) ---------------------------------------
 * blob_host: fp32, in this order
 *   mean[3], std[3]                                        (Encoder.mean / Encoder.std)
 *   stem weight [64][7][7][4]  (cout, kh, kw, cin; cin 0..2 = Encoder.conv1, cin 3 = Encoder.conv1_p)
 *   stem BN: weight[64], bias[64], running_mean[64], running_var[64]      (Encoder.bn1)
 *   for each of the 52 convs of res2..res5 in execution order (per block: conv1, conv2,
 *   [downsample.0 if first block of the stage], conv3):
 *       weight [cout][kh][kw][cin], BN weight[cout], bias[cout], running_mean[cout], running_var[cout]
 *   fc1.weight[2048], fc1.bias[1]
 * ivosw_assess_blob_floats() returns the expected length. */
IVOSW_API size_t ivosw_assess_blob_floats(void);
IVOSW_API int ivosw_assess_load(ivosw_ctx* ctx, const float* blob_host, size_t n_floats);

/* AssessNet.forward (models/assessment.py:164-182).
 * frames_dev: B x 3 x H x W fp32 NCHW, frame b at frames_dev + b*frame_stride (floats);
 * prob_dev:   B planes of H x W fp32, plane b at prob_dev + b*prob_stride (floats) — so
 *             all_P[:, i+1] of a T x (O+1) x H x W tensor is passed without a copy;
 * score_dev:  B fp32.  boxes_dev (nullable): B x 4 fp32 [yc, xc, h, w] of all2yxhw (:110-161). */
IVOSW_API int ivosw_assess_forward(ivosw_ctx* ctx, const float* frames_dev, long long frame_stride,
                         const float* prob_dev, long long prob_stride, int B, int H, int W,
                         float* score_dev, float* boxes_dev, void* stream);

/* Parity probes (tests only).  After ivosw_enable_probes(ctx, 1) every ivosw_assess_forward
 * keeps the intermediates of its LAST chunk; ivosw_assess_probe converts one of them to NCHW
 * fp32 in out_dev.  which: 0 ROI crop (4 x 256 x 256: normalised RGB + prob), 1 after maxpool
 * (64 x 64 x 64), 2..5 r2..r5.  dims4_out receives {n, c, h, w}. */
IVOSW_API int ivosw_enable_probes(ivosw_ctx* ctx, int enable);
IVOSW_API int ivosw_assess_probe(ivosw_ctx* ctx, int which, float* out_dev, size_t capacity_floats,
                       int* dims4_out, void* stream);

/* ---- one scoring round (utils/utils_agent.py::recommend_frame, wild/ours, :111-122) -------
 * frames_dev: T x 3 x H x W; probs_dev: T x (O+1) x H x W (channel 0 = background, unused);
 * annotated_counts_host: T doubles (histogram of annotated frames, :112-113).
 * Scores frames [t_begin, t_end) for every object, averages over objects in fp64 (:120),
 * and — when the range covers the whole clip — runs Brain and the argmax on device.
 * Outputs (host, written after an internal stream synchronise):
 *   mask_quality_host[t_end - t_begin] doubles, scores_host (nullable) (t_end-t_begin) x O fp32,
 *   q_host (nullable) T fp32, next_frame (nullable) int.
 * q_host / next_frame are only produced when t_begin == 0 && t_end == T. */
IVOSW_API int ivosw_round_device(ivosw_ctx* ctx, const float* frames_dev, const float* probs_dev,
                       int T, int O, int H, int W, int t_begin, int t_end,
                       const double* annotated_counts_host, double* mask_quality_host,
                       float* scores_host, float* q_host, int* next_frame, void* stream);

/* Same round from HOST buffers (the end-to-end path bench.py reports as `e2e`): the
 * host->device copies of frames and probabilities are issued inside the call, in
 * frame chunks that overlap with the scoring of the previous chunk. */
IVOSW_API int ivosw_round_host(ivosw_ctx* ctx, const float* frames_host, const float* probs_host,
                     int T, int O, int H, int W,
                     const double* annotated_counts_host, double* mask_quality_host,
                     float* scores_host, float* q_host, int* next_frame, void* stream);

/* Brain + argmax on an already gathered quality vector (multi-GPU: after the all-gather).
 * mask_quality_host: T doubles; annotated_counts_host: T doubles. Synchronises. */
IVOSW_API int ivosw_agent_action(ivosw_ctx* ctx, const double* mask_quality_host,
                       const double* annotated_counts_host, int T,
                       float* q_host, int* next_frame, void* stream);

/* ---- frame-sharded round (multi-GPU) -----------------------------------
 * ivosw_score_shard: the scoring half of the round for frames [t_begin, t_end), fully
 * asynchronous: writes the float64 per-frame mean quality to mq_dev (device, t_end - t_begin
 * doubles — e.g. this rank's slice of the all-gather buffer) and, if scores_dev is not NULL,
 * the per-object scores as [O][t_end - t_begin] fp32.  No host synchronisation.
 * ivosw_agent_action_dev: Brain + argmax on the gathered device vector mq_dev[T]
 * (annotated_counts_host: T doubles); synchronises the stream and returns q / index on the host. */
IVOSW_API int ivosw_score_shard(ivosw_ctx* ctx, const float* frames_dev, const float* probs_dev,
                      int T, int O, int H, int W, int t_begin, int t_end,
                      double* mq_dev, float* scores_dev, void* stream);
/* ivosw_score_shard_host: as ivosw_score_shard, with the clip in HOST memory (full-clip base pointers;
 * only frames [t_begin, t_end) are read): chunked upload overlapped with scoring.  Asynchronous. */
IVOSW_API int ivosw_score_shard_host(ivosw_ctx* ctx, const float* frames_host, const float* probs_host,
                           int T, int O, int H, int W, int t_begin, int t_end, double* mq_dev, void* stream);
IVOSW_API int ivosw_agent_action_dev(ivosw_ctx* ctx, const double* mq_dev, const double* annotated_counts_host,
                           int T, float* q_host, int* next_frame, void* stream);

/* ---- per-stage device timing (bench.py's roofline leg) ----------------------------------------
 * While enabled, CUDA events are recorded on the launching stream around each stage of every
 * chunk; ivosw_stage_times synchronises them and returns the accumulated milliseconds since the
 * last reset: [0] bbox+ROI crop, [1] stem conv+maxpool, [2] res2..res5 conv stack, [3] pool+FC,
 * [4] Brain (3 launches), and the number of conv-stack launches in n_conv_launches. */
/* ---- fp16 range guard of the split-fp16 encoder ------------------------------------------------
 * Activations are carried as two fp16 planes; a value beyond +-65504 is clamped.  The epilogues count
 * the (thread, tile) pairs in which that happened; this call synchronises `stream`, returns the count
 * since the last reset and, when it is non-zero, leaves a warning in ivosw_last_error().  0 for every
 * network whose activations stay inside the fp16 range (the reference's trained weights: O(1..100)). */
IVOSW_API int ivosw_conv_saturation_count(ivosw_ctx* ctx, long long* count_out, int reset, void* stream);

#define IVOSW_NUM_STAGES 5
IVOSW_API int ivosw_stage_timing(ivosw_ctx* ctx, int enable);
IVOSW_API int ivosw_stage_times(ivosw_ctx* ctx, float* ms_out /*[IVOSW_NUM_STAGES]*/, long long* n_conv_launches,
                      int reset);

/* ---- test hook: one bottleneck convolution in isolation -------------------------------------
 * Runs layer `layer_index` (0..51, execution order of res2..res5) on in_dev (B x H x W x Cin fp32
 * NHWC) with optional residual_dev (B x OH x OW x Cout fp32 NHWC) into out_dev (fp32 NHWC), with
 * its folded BatchNorm and ReLU, using `conv_mode`.  The tensor-core modes convert to and from the
 * split-fp16 planes internally.  Lets tests compare the tcgen05 kernel with the fp32 CUDA-core
 * kernel layer by layer.  dims_out (nullable) receives {cin, in_hw, cout, out_hw, k, stride}. */
IVOSW_API int ivosw_debug_conv(ivosw_ctx* ctx, int layer_index, int conv_mode, const float* in_dev,
                     const float* residual_dev, float* out_dev, int B, int* dims_out, void* stream);

/* ---- MANet round tail (utils/utils_manet.py:76-81,109-114,146-150,160-161) ---------------
 * logits_dev: T x C x h x w fp32 (C = O+1).  Bilinear upsample (align_corners=True) to
 * H x W, per-pixel first-max argmax -> masks_dev (T x H x W fp32, nullable), channel
 * softmax -> all_p_dev (T x C x H x W fp32, nullable). */
IVOSW_API int ivosw_manet_tail(ivosw_ctx* ctx, const float* logits_dev, int T, int C, int h, int w,
                     int H, int W, float* masks_dev, float* all_p_dev, void* stream);

/* utils/utils_manet.py::rough_ROI (22-39): labels_dev / out_dev B x 1 x h x w fp32 (may alias); keeps the labels
 * inside the +-dist bounding box of the pixels != -1 and writes 0 elsewhere.  Synchronises; returns
 * IVOSW_ERR_INVALID if an image has no such pixel (the reference raises there). */
IVOSW_API int ivosw_rough_roi(ivosw_ctx* ctx, const float* labels_dev, float* out_dev, int B, int h, int w,
                    int dist, void* stream);

/* ---- MANet feature extractor (IntVOS.extract_feature, call site eval_agent_manet.py:316-328) --------------------
 * DeepLabv3+ ResNet-101 (output stride 16) + ASPP + shortcut decoder + semantic-embedding head: frames B x 3 x H x W
 * (normalised, device) -> embedding B x 100 x h/4 x w/4 fp32 (device), e.g. 480 x 854 -> 120 x 214.
 * RESTATEMENT: the network's source is not part of the reference tree; the architecture and the parameter blob follow
 * ivosw/manet_arch.py (hyper-parameters of utils/config_manet/config.py:108-120) and parity is claimed against
 * oracle/manet_encoder_ref.py only (parity unpinned).  Blob: per convolution in manet_arch.convs()
 * order, weight as OHWI then BatchNorm gamma, beta, running_mean (minus the conv bias where there is one), running_var. */
IVOSW_API size_t ivosw_manet_encoder_blob_floats(void);
IVOSW_API int ivosw_manet_encoder_load(ivosw_ctx* ctx, const float* blob_host, size_t n_floats);
IVOSW_API int ivosw_manet_encoder_forward(ivosw_ctx* ctx, const float* frames_dev, int B, int H, int W, float* embedding_dev,
                                          void* stream);

/* ---- AssessNet optimisation step (quality_assessment.py::train :240-269; BASELINE config C5) ---------------------
 * ivosw_assess_train_begin   (re)starts training from a parameter blob in ivosw_assess_load's layout: parameters and
 *                            BatchNorm buffers go to the device, gradients and SGD momentum buffers start at zero.
 * ivosw_assess_train_step    one loop iteration on B samples (frames B x 3 x H x W, probabilities B x H x W, device; targets
 *                            B fp32 = the J&F metric; valid B int32 = `union[n] > 0`): train-mode forward (batch statistics,
 *                            running statistics updated), masked MSE, backward, gradients ACCUMULATED onto the previous
 *                            steps' (the reference never zeroes them) and clamped to [-1, 1], SGD with momentum and weight
 *                            decay (apply_update = 0: forward + backward only; the gradient of this step stays in the buffer
 *                            ivosw_assess_train_grads returns — all-reduce it across ranks, then ivosw_assess_train_apply).  loss_host: the loss (NaN when no sample is valid: the reference
 *                            `continue`s before backward); pred_host (nullable): B predictions.  Synchronises.
 * ivosw_assess_train_export  parameters + BatchNorm buffers (blob layout, loads back with ivosw_assess_load) and / or the
 *                            accumulated clamped gradients (same layout; non-parameter slots zero) to host memory.
 * ivosw_assess_train_grads   device pointer to the CURRENT step's raw gradient (blob layout) for an all-reduce between
 *                            ivosw_assess_train_step(apply_update = 0) calls of several ranks. */
IVOSW_API int ivosw_assess_train_begin(ivosw_ctx* ctx, const float* blob_host, size_t n_floats);
IVOSW_API int ivosw_assess_train_step(ivosw_ctx* ctx, const float* frames_dev, long long frame_stride_floats,
                                      const float* prob_dev, long long prob_stride_floats, int B, int H, int W,
                                      const float* targets_dev, const int* valid_dev, float lr, float momentum,
                                      float weight_decay, int apply_update, float* loss_host, float* pred_host, void* stream);
IVOSW_API int ivosw_assess_train_apply(ivosw_ctx* ctx, float lr, float momentum, float weight_decay, void* stream);
IVOSW_API int ivosw_assess_train_export(ivosw_ctx* ctx, float* blob_host, float* grad_host, void* stream);
IVOSW_API int ivosw_assess_train_grads(ivosw_ctx* ctx, float** grads_dev, size_t* n_floats);

/* ---- the exchange step of the frame-sharded round over NVLink peer memory ----------------
 * Instead of an NCCL all-gather of ceil(T / G) doubles per rank, every rank writes its slice of the per-frame quality
 * vector straight into every rank's gather buffer (peer stores) and raises a flag there; Brain starts once all G flags
 * of the local buffer carry the round's number.  Both are small kernels of this library on the caller's stream: no host
 * synchronisation, no collective call, CUDA-graph capturable (csrc/gather.cu).
 *   ivosw_gather_create  allocates this rank's buffer (capacity = largest T) and returns its 64-byte cudaIpcMemHandle_t
 *   ivosw_gather_open    handles: world x 64 bytes, rank order (exchanged by the host framework, e.g. one all_gather at
 *                        start-up); maps every peer buffer (cudaIpcOpenMemHandle, peer access enabled lazily)
 *   ivosw_gather_post    after ivosw_score_shard(.., mq_dev = mq_local_dev ..): posts frames [offset, offset + n_local) of this
 *                        round; EVERY rank calls it exactly once per round (n_local = 0 for an empty shard)
 *   ivosw_agent_action_gathered  waits for all ranks' slices of the round, then Brain + argmax (as ivosw_agent_action_dev);
 *                        mask_quality_host (nullable) receives the gathered T doubles.  Synchronises the stream. */
IVOSW_API int ivosw_gather_create(ivosw_ctx* ctx, int world, int rank, int capacity, void* handle_out /*64 bytes*/);
IVOSW_API int ivosw_gather_open(ivosw_ctx* ctx, const void* handles /*world x 64 bytes*/);
IVOSW_API int ivosw_gather_post(ivosw_ctx* ctx, const double* mq_local_dev, int n_local, int offset, void* stream);
IVOSW_API int ivosw_agent_action_gathered(ivosw_ctx* ctx, const double* annotated_counts_host, int T, float* q_host,
                                          int* next_frame, double* mask_quality_host, void* stream);

/* ---- ATNet round wrapper glue (utils/utils_atnet.py::run_VOS_singleiact; config C3) -------------
 * The ATNet networks are external (yuk6heo/IVOS-ATNet, not in the reference tree); these are the wrapper's own
 * element-wise passes, each one kernel:
 *   reflect_pad    :95-96   torch.nn.ReflectionPad2d((left, right, top, bottom)) on `planes` h x w planes
 *   sigmoid_blend  :101-102,124-126,146-150  prob = sigmoid(logit); blended = alpha*prob + one_minus_alpha*prev
 *                  (prev_dev NULL: blended = prob, the annotated frame); blended_dev may alias prev_dev
 *   assemble       :157-159 all_P[T][O+1][H][W]: channel 0 zero, channels 1..O = prob_map[T][O][PH][PW] cropped at
 *                  (y0, x0) — the contiguous equivalent of cat([zeros, prob_map], 1)[:, :, hpad1:-hpad2, wpad1:-wpad2] */
IVOSW_API int ivosw_atnet_reflect_pad(ivosw_ctx* ctx, const float* in_dev, float* out_dev, int planes, int h, int w,
                                      int left, int right, int top, int bottom, void* stream);
IVOSW_API int ivosw_atnet_sigmoid_blend(ivosw_ctx* ctx, const float* logit_dev, const float* prev_dev, float* prob_dev,
                                        float* blended_dev, long long n, float alpha, float one_minus_alpha, void* stream);
IVOSW_API int ivosw_atnet_assemble(ivosw_ctx* ctx, const float* prob_map_dev, float* all_p_dev, int T, int O, int PH, int PW,
                                   int y0, int x0, int H, int W, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* IVOSW_B200_H */

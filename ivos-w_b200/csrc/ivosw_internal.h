// Internal declarations shared by the CUDA translation units behind include/ivosw_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <vector>

#include "../../include/ivosw_b200.h"

namespace ivosw {

constexpr int ROI = 256;           // models/assessment.py:170 dst_size
constexpr int CROP_PH = 262;       // zero-bordered crop canvas of the tensor-core stem: image at (3, 3),
constexpr int CROP_PW = 264;       //   rows -3..258, columns -3..260 of the 7x7 stride-2 pad-3 convolution
constexpr float BN_EPS = 1e-5f;    // torch BatchNorm2d default
constexpr int BRAIN_H = 128;       // models/agent.py:14 hidden_channels
constexpr int BRAIN_G = 4 * BRAIN_H;

void set_error(const std::string& msg);
int cuda_fail(cudaError_t e, const char* what, const char* file, int line);

#define IVOSW_CUDA(call)                                                         \
    do {                                                                         \
        cudaError_t _e = (call);                                                 \
        if (_e != cudaSuccess) return ::ivosw::cuda_fail(_e, #call, __FILE__, __LINE__); \
    } while (0)

#define IVOSW_REQUIRE(cond, msg)                                  \
    do {                                                          \
        if (!(cond)) {                                            \
            ::ivosw::set_error(std::string("invalid argument: ") + (msg)); \
            return IVOSW_ERR_INVALID;                             \
        }                                                         \
    } while (0)

// One convolution of res2..res5 (mirrors ivosw/arch.py::resnet50_convs).
struct ConvLayer {
    int cin, cout, k, stride, pad;
    int in_hw, out_hw;
    bool relu;
    int residual;         // 0 none, 1 identity (block input), 2 downsample output
    bool is_downsample;
    bool first_of_block;  // conv1 of a bottleneck: its input is the block input
    // device parameters
    float* w_f32 = nullptr;    // [cout][k][k][cin]
    float* scale = nullptr;    // gamma / sqrt(var + eps)
    float* shift = nullptr;    // beta - mean * scale
    __half* w_hi = nullptr;    // tensor-core path: fp16 head of the weight, [cout][k*k*cin]
    __half* w_lo = nullptr;    //                   fp16((w - hi) * 2048)
};

std::vector<ConvLayer> make_resnet50_layers();

// One tensor-core convolution on power-of-two canvases (the VOS encoder: manet_encoder.cu / conv_tc.cu launch_conv_tc_g)
struct GConv {
    int cin = 0, cout = 0;                      // multiples of 64 (zero-padded channels)
    int k = 1, stride = 1, dil = 1;
    int in_hp = 0, in_wp = 0, out_hp = 0, out_wp = 0;   // canvases (powers of two)
    int valid_h = 0, valid_w = 0;               // feature-map size inside the output canvas
    int relu = 1;
    __half* w_hi = nullptr;                     // [cout][k*k*cin + cin2], BatchNorm scale folded in when fused
    __half* w_lo = nullptr;
    float* scale = nullptr;
    float* shift = nullptr;
    int cin2 = 0, stride2 = 1, in2_hp = 0, in2_wp = 0;   // fused 1x1 second GEMM (downsample branch)
};

// conv3 + downsample of a stage's first bottleneck as one GEMM (conv_tc.cu: launch_conv_tc with `fuse`): the concatenated,
// BatchNorm-scaled weight [cout][cin3 + cin_d] as split-fp16 planes, scale = 1, shift = shift3 + shift_d.
struct FusedTail {
    __half* w_hi = nullptr;
    __half* w_lo = nullptr;
    float* scale = nullptr;
    float* shift = nullptr;
    int k_total = 0;
    int cin2 = 0, in_hw2 = 0, stride2 = 1;      // the downsample branch's input geometry
};

// Where the inputs of scoring unit u = (object o, frame f) live: u = u0 + local index, f = u % nF,
// o = u / nF.  One AssessNet.forward call is nF = B, obj_stride = 0; a whole round is nF = frames of
// the shard with obj_stride = H*W walking over all_P[:, 1:].
struct UnitAddr {
    const float* frames; long long frame_stride;
    const float* prob; long long prob_stride; long long obj_stride;
    int nF; int u0;
    __host__ __device__ const float* frame(int local) const {
        int u = u0 + local; return frames + (long long)(u % nF) * frame_stride;
    }
    __host__ __device__ const float* prob_plane(int local) const {
        int u = u0 + local; return prob + (long long)(u % nF) * prob_stride + (long long)(u / nF) * obj_stride;
    }
};

struct DeviceBuffer {
    void* p = nullptr;
    size_t bytes = 0;
};

// An activation tensor in split-fp16 form (conv_tc.cu): x = hi + lo / 2048, NHWC, two planes.
struct SplitAct {
    __half* hi;
    __half* lo;
};
// The two planes share one DeviceBuffer sized for the fp32 tensor: hi in the first half, lo in the second.
inline SplitAct split_view(const DeviceBuffer& b) {
    return SplitAct{reinterpret_cast<__half*>(b.p),
                    reinterpret_cast<__half*>(reinterpret_cast<char*>(b.p) + b.bytes / 2)};
}

}  // namespace ivosw

struct ivosw_ctx {
    int device = 0;
    int conv_mode = 0;
    int sm_count = 148;
    long long launches = 0;

    // Brain
    bool brain_loaded = false;
    float* brain_params = nullptr;   // canonical order, see header
    float* brain_d1t = nullptr;      // decoder_fc1.weight transposed [256][128]
    unsigned int* brain_done_count = nullptr;   // per-sequence completion counters of the decoder groups (zero at rest)
    float* brain_whh_t = nullptr;    // weight_hh transposed/packed for the recurrent kernel
    ivosw::DeviceBuffer brain_gi, brain_h, brain_state, brain_q, brain_arg;
    // Double-DQN training step (dqn.cu): target network, Adam moments, saved activations
    bool target_loaded = false;
    float* target_params = nullptr;
    float* target_whh_t = nullptr;
    float* target_d1t = nullptr;
    float* adam_m = nullptr;
    float* adam_v = nullptr;
    long long adam_step = 0;
    ivosw::DeviceBuffer dqn_ws;

    // AssessNet
    bool assess_loaded = false;
    float* stem_w = nullptr;         // [196][64]: (kh,kw,cin) x cout
    float* stem_scale = nullptr;
    float* stem_shift = nullptr;
    void* stem_wpack = nullptr;      // stem weight as the smem image of stem_tc.cu (fp16 hi/lo planes)
    float mean[3], stdv[3];
    float* fc_w = nullptr;           // 2048
    float fc_b = 0.f;
    std::vector<ivosw::ConvLayer> layers;
    ivosw::FusedTail fused_tail[4];         // res2.0 .. res5.0: conv3 + downsample as one GEMM
    bool fuse_ds = true;                    // IVOSW_FUSE_DS=0: two kernels and a round trip of the downsample output
    // fp16 range guard: epilogue threads whose tile produced a value clamped to +-65504 (split-fp16 planes cannot hold
    // more) bump this device counter; ivosw_conv_saturation_count reads it.  0 on every in-range network.
    unsigned long long* sat_count = nullptr;

    // workspace (grown on demand, per chunk of `chunk_cap` samples)
    int chunk_cap = 0;
    ivosw::DeviceBuffer crop_hi, crop_lo;   // tensor-core path: padded split-fp16 crop planes (roi.cu)
    ivosw::DeviceBuffer bbox_min, bbox_max, boxes, crop, c1, pool, actX, actY, actDS, actT1, actT2, scores, mq;
    // what the probe entry point can read back (valid for the last chunk processed)
    bool probes_on = false;
    int last_chunk_b = 0;
    ivosw::DeviceBuffer probe_buf[6];   // fp32 NHWC copies: crop, pool, r2, r3, r4, r5

    // per-stage timing (ivosw_stage_timing)
    bool timing_on = false;
    struct StageEvt { int stage; cudaEvent_t a, b; };
    std::vector<StageEvt> stage_evts;
    std::vector<cudaEvent_t> evt_pool;
    float stage_ms[IVOSW_NUM_STAGES] = {0, 0, 0, 0, 0};
    long long conv_launches_timed = 0;

    // CUDA-graph replay of whole rounds / shards (capi.cu: run_graphed)
    struct GraphKey {
        const void* p0; const void* p1; const void* p2;
        int T, O, H, W, tb, te, mode, kind, flags;
    };
    struct GraphEntry {
        GraphKey key;
        unsigned long long epoch = 0;
        int seen = 0;
        cudaGraphExec_t exec = nullptr;
        std::vector<StageEvt> evts;          // external event-record nodes baked into the graph (timing on)
        long long launches = 0, conv_launches = 0;
    };
    std::vector<GraphEntry> graphs;
    bool graphs_on = true;
    bool capturing = false;
    std::vector<StageEvt>* capture_evts = nullptr;
    GraphEntry* last_graph = nullptr;        // replayed graph whose timing events still have to be read
    cudaStream_t graph_stream = nullptr;
    cudaEvent_t g_ev1 = nullptr, g_ev2 = nullptr;

    // persistent conv-stack kernel (conv_stack.cu): one output buffer per layer, cached plans
    ivosw::DeviceBuffer stack_arena;
    int stack_arena_cap = 0;
    int chunk_cap_seen = 0;
    void* stack_state = nullptr;
    void* enc_state = nullptr;               // manet_encoder.cu
    void* train_state = nullptr;             // train.cu
    void* gather_state = nullptr;            // gather.cu: peer-memory exchange of the frame-sharded round
    bool stack_on = false;
    unsigned long long assess_version = 0;   // bumped by every ivosw_assess_load

    // host-staged rounds
    ivosw::DeviceBuffer stage_frames, stage_probs, scores_all;
    std::vector<cudaEvent_t> chunk_evts;
    void* pinned_small = nullptr;    // small pinned scratch for D2H results
    size_t pinned_small_bytes = 0;
    cudaStream_t copy_stream = nullptr;
    cudaStream_t aux_stream = nullptr;     // row-range pre-pass of the host-buffer path
    ivosw::DeviceBuffer band_min, band_max, band_rows;
    int4* pinned_rows = nullptr; size_t pinned_rows_n = 0;
    long long last_h2d_bytes = 0;          // bytes the last host-buffer call actually sent
};

namespace ivosw {

int ensure(DeviceBuffer& b, size_t bytes);
int ensure_pinned(ivosw_ctx* c, size_t bytes);
// RAII-free stage bracket: stage_begin returns an index (or -1 when timing is off)
int stage_begin(ivosw_ctx* c, int stage, cudaStream_t s);
void stage_end(ivosw_ctx* c, int idx, cudaStream_t s);
void release(DeviceBuffer& b);

// ---- roi.cu
int launch_bbox(ivosw_ctx* c, const UnitAddr& ua, int B, int H, int W, cudaStream_t s, DeviceBuffer* mn_buf = nullptr,
                DeviceBuffer* mx_buf = nullptr);
int launch_roi_rows(ivosw_ctx* c, const int2* mn, const int2* mx, int nF, int O, int H, int W, int4* rows, cudaStream_t s);
int launch_roi_sample(ivosw_ctx* c, const UnitAddr& ua, int B, int H, int W, float* boxes_out, bool split,
                      cudaStream_t s);
int launch_crop_merge(ivosw_ctx* c, float* out, int B, int use_lo, cudaStream_t s);
// ---- stem.cu
int launch_stem(ivosw_ctx* c, int B, cudaStream_t s);
// ---- stem_tc.cu
int stem_tc_pack(ivosw_ctx* c, const float* w_ohwi);
int launch_stem_tc(ivosw_ctx* c, int B, const SplitAct& out, int terms, cudaStream_t s);
// ---- conv_simt.cu
int launch_conv_simt(ivosw_ctx* c, const ConvLayer& L, const float* in, const float* residual, float* out,
                     int B, cudaStream_t s);
// ---- conv_tc.cu
int launch_conv_tc(ivosw_ctx* c, const ConvLayer& L, const SplitAct& in, const SplitAct* residual, const SplitAct& out,
                   int B, int terms, cudaStream_t s, const FusedTail* fuse = nullptr, const SplitAct* in2 = nullptr);
int launch_conv_tc_g(ivosw_ctx* c, const GConv& L, const SplitAct& in, const SplitAct* in2, const SplitAct* residual,
                     const SplitAct& out, int out_ld, int B, int terms, cudaStream_t s);
int launch_split(ivosw_ctx* c, const float* in, const SplitAct& out, long long n, cudaStream_t s);
int launch_merge(ivosw_ctx* c, const SplitAct& in, float* out, long long n, int use_lo, cudaStream_t s);
// ---- manet_encoder.cu (MANet feature extractor, restatement)
size_t manet_encoder_blob_floats();
int manet_encoder_load(ivosw_ctx* c, const float* blob, size_t n_floats);
int manet_encoder_forward(ivosw_ctx* c, const float* frames, int B, int H, int W, float* out, int terms, cudaStream_t s);
void manet_encoder_release(ivosw_ctx* c);
// ---- train.cu (AssessNet optimisation step, config C5)
int train_begin(ivosw_ctx* c, const float* blob, size_t n_floats);
int train_step(ivosw_ctx* c, const UnitAddr& ua, int B, int H, int W, const float* targets_dev, const int* valid_dev, float lr,
               float momentum, float wd, int apply, float* loss_host, float* pred_host, cudaStream_t s);
int train_apply(ivosw_ctx* c, float lr, float momentum, float wd, cudaStream_t s);
int train_export(ivosw_ctx* c, float* blob_host, float* grad_host, cudaStream_t s);
int train_grad_buffer(ivosw_ctx* c, float** gnew_dev, size_t* n);
void train_release(ivosw_ctx* c);
// ---- gather.cu
int gather_create(ivosw_ctx* c, int world, int rank, int cap, void* handle_out);
int gather_open(ivosw_ctx* c, const void* handles);
int gather_post(ivosw_ctx* c, const double* mq_local_dev, int n_local, int offset, cudaStream_t s);
int gather_wait_pack(ivosw_ctx* c, const double* ann_dev, int T, float* state_dev, double* mq_out_dev, cudaStream_t s);
void gather_release(ivosw_ctx* c);
// ---- conv_stack.cu
int launch_conv_stack(ivosw_ctx* c, const SplitAct& in, int B, int terms, cudaStream_t s, SplitAct* out, SplitAct* stage_out);
void conv_stack_release(ivosw_ctx* c);
unsigned long long current_alloc_epoch();
// ---- head.cu
int launch_gap_fc(ivosw_ctx* c, const float* r5, int B, float* scores, cudaStream_t s);
int launch_gap_fc_split(ivosw_ctx* c, const SplitAct& r5, int use_lo, int B, float* scores, cudaStream_t s);
int launch_object_mean(ivosw_ctx* c, const float* scores, int T, int O, const double* ann_dev, double* mq_dev,
                       float* state_dev, cudaStream_t s);
int launch_pack_state(ivosw_ctx* c, const double* mq_dev, const double* ann_dev, int T, float* state_dev,
                      cudaStream_t s);
// ---- brain.cu
int brain_pack(ivosw_ctx* c);
int launch_brain(ivosw_ctx* c, const float* state, int N, int T, float* q, int* argmax, cudaStream_t s);
struct BrainSaves { float *A1, *E, *G, *C, *HP, *H; };   // activations kept for the DQN backward pass
int launch_brain_ex(ivosw_ctx* c, const float* params, const float* whh_pack, const float* d1t, const float* state,
                    int N, int T, float* q, int* argmax, const BrainSaves* sv, cudaStream_t s);
int brain_pack_into(ivosw_ctx* c, const float* params, float* whh_pack, float* d1t, cudaStream_t s);
// ---- dqn.cu
int dqn_update(ivosw_ctx* c, const float* state, const float* new_state, const int* action, const float* reward_step,
               const float* reward_done, int N, int T, float gamma, float lr, float weight_decay, float* loss_host,
               float* grads_out_dev, bool apply, cudaStream_t s);
int dqn_apply(ivosw_ctx* c, float* grad, float lr, float weight_decay, cudaStream_t s);
// ---- manet_tail.cu
int launch_manet_tail(ivosw_ctx* c, const float* logits, int T, int C, int h, int w, int H, int W, float* masks,
                      float* all_p, cudaStream_t s);
int launch_rough_roi(ivosw_ctx* c, const float* in, float* out, int B, int h, int w, int dist, int* empty_flag_dev,
                     cudaStream_t s);
// ---- atnet_glue.cu
int launch_reflect_pad(ivosw_ctx* c, const float* in, float* out, long long planes, int h, int w, int left, int right, int top,
                       int bottom, cudaStream_t s);
int launch_sigmoid_blend(ivosw_ctx* c, const float* logit, const float* prev, float* prob, float* blended, long long n,
                         float alpha, float beta, cudaStream_t s);
int launch_atnet_assemble(ivosw_ctx* c, const float* prob_map, float* all_p, int T, int O, int PH, int PW, int y0, int x0,
                          int H, int W, cudaStream_t s);
// ---- probe helper (NHWC -> NCHW)
int launch_nhwc_to_nchw(ivosw_ctx* c, const float* in, float* out, int B, int HW, int C, cudaStream_t s);

}  // namespace ivosw

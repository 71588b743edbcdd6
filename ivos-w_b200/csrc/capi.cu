// C-ABI entry points (include/ivosw_b200.h): context, weight loading, orchestration of the kernels.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "ivosw_internal.h"

namespace ivosw {

static thread_local std::string g_err;

void set_error(const std::string& msg) { g_err = msg; }

int cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
    char buf[512];
    if (e == cudaErrorMemoryAllocation) {
        snprintf(buf, sizeof buf, "CUDA out of memory (%s at %s:%d)", what, file, line);
        g_err = buf;
        cudaGetLastError();
        return IVOSW_ERR_OOM;
    }
    snprintf(buf, sizeof buf, "CUDA error %d (%s): %s at %s:%d", (int)e, cudaGetErrorString(e), what, file, line);
    g_err = buf;
    cudaGetLastError();
    return IVOSW_ERR_CUDA;
}

// bumped whenever a workspace buffer moves: captured CUDA graphs hold raw pointers and must be rebuilt
static unsigned long long g_alloc_epoch = 1;
unsigned long long current_alloc_epoch() { return g_alloc_epoch; }

int ensure(DeviceBuffer& b, size_t bytes) {
    if (b.bytes >= bytes && b.p) return IVOSW_OK;
    ++g_alloc_epoch;
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.bytes = 0; }
    size_t want = (bytes + 255) & ~(size_t)255;
    IVOSW_CUDA(cudaMalloc(&b.p, want));
    b.bytes = want;
    return IVOSW_OK;
}

void release(DeviceBuffer& b) {
    if (b.p) cudaFree(b.p);
    b.p = nullptr; b.bytes = 0;
}

static cudaEvent_t take_event(ivosw_ctx* c) {
    cudaEvent_t e = nullptr;
    if (!c->evt_pool.empty()) { e = c->evt_pool.back(); c->evt_pool.pop_back(); return e; }
    cudaEventCreate(&e);
    return e;
}

int stage_begin(ivosw_ctx* c, int stage, cudaStream_t s) {
    if (!c->timing_on) return -1;
    ivosw_ctx::StageEvt ev{stage, take_event(c), take_event(c)};
    if (c->capturing) {   // external record node: the event is re-recorded by every replay and can be timed
        cudaEventRecordWithFlags(ev.a, s, cudaEventRecordExternal);
        c->capture_evts->push_back(ev);
        return (int)c->capture_evts->size() - 1;
    }
    cudaEventRecord(ev.a, s);
    c->stage_evts.push_back(ev);
    return (int)c->stage_evts.size() - 1;
}

void stage_end(ivosw_ctx* c, int idx, cudaStream_t s) {
    if (idx < 0) return;
    if (c->capturing) cudaEventRecordWithFlags((*c->capture_evts)[idx].b, s, cudaEventRecordExternal);
    else cudaEventRecord(c->stage_evts[idx].b, s);
}

// timing events of the graph replayed last (must be read before that graph is launched again)
static void drain_graph_events(ivosw_ctx* c) {
    if (!c->last_graph) return;
    for (auto& ev : c->last_graph->evts) {
        float ms = 0.f;
        if (cudaEventSynchronize(ev.b) == cudaSuccess && cudaEventElapsedTime(&ms, ev.a, ev.b) == cudaSuccess)
            c->stage_ms[ev.stage] += ms;
    }
    c->last_graph = nullptr;
}

static void drain_stage_events(ivosw_ctx* c) {
    drain_graph_events(c);
    for (auto& ev : c->stage_evts) {
        float ms = 0.f;
        if (cudaEventSynchronize(ev.b) == cudaSuccess && cudaEventElapsedTime(&ms, ev.a, ev.b) == cudaSuccess)
            c->stage_ms[ev.stage] += ms;
        c->evt_pool.push_back(ev.a);
        c->evt_pool.push_back(ev.b);
    }
    c->stage_evts.clear();
}

std::vector<ConvLayer> make_resnet50_layers() {
    // mirrors ivosw/arch.py::resnet50_convs — torchvision Bottleneck, stride on the 3x3 conv
    std::vector<ConvLayer> v;
    const int planes_[4] = {64, 128, 256, 512}, blocks_[4] = {3, 4, 6, 3}, stride_[4] = {1, 2, 2, 2};
    int inplanes = 64, hw = 64;
    for (int st = 0; st < 4; ++st) {
        for (int b = 0; b < blocks_[st]; ++b) {
            const int s = b == 0 ? stride_[st] : 1, planes = planes_[st], out_hw = hw / s;
            ConvLayer c1{}; c1.cin = inplanes; c1.cout = planes; c1.k = 1; c1.stride = 1; c1.pad = 0;
            c1.in_hw = hw; c1.out_hw = hw; c1.relu = true; c1.residual = 0; c1.first_of_block = true;
            v.push_back(c1);
            ConvLayer c2{}; c2.cin = planes; c2.cout = planes; c2.k = 3; c2.stride = s; c2.pad = 1;
            c2.in_hw = hw; c2.out_hw = out_hw; c2.relu = true;
            v.push_back(c2);
            if (b == 0) {
                ConvLayer d{}; d.cin = inplanes; d.cout = planes * 4; d.k = 1; d.stride = s; d.pad = 0;
                d.in_hw = hw; d.out_hw = out_hw; d.relu = false; d.is_downsample = true;
                v.push_back(d);
            }
            ConvLayer c3{}; c3.cin = planes; c3.cout = planes * 4; c3.k = 1; c3.stride = 1; c3.pad = 0;
            c3.in_hw = out_hw; c3.out_hw = out_hw; c3.relu = true; c3.residual = b == 0 ? 2 : 1;
            v.push_back(c3);
            inplanes = planes * 4;
            hw = out_hw;
        }
    }
    return v;
}

static size_t assess_blob_floats() {
    size_t n = 6 + (size_t)64 * 49 * 4 + 4 * 64;
    for (const ConvLayer& L : make_resnet50_layers()) n += (size_t)L.cout * L.k * L.k * L.cin + 4 * (size_t)L.cout;
    return n + 2048 + 1;
}

static int upload(float** dst, const float* src, size_t n) {
    if (!*dst) IVOSW_CUDA(cudaMalloc(dst, n * sizeof(float)));
    IVOSW_CUDA(cudaMemcpy(*dst, src, n * sizeof(float), cudaMemcpyHostToDevice));
    return IVOSW_OK;
}

// BatchNorm (eval) -> per-channel scale / shift, evaluated in double then rounded once.
static void fold_bn(const float* g, const float* b, const float* m, const float* v, int n, std::vector<float>& sc,
                    std::vector<float>& sh) {
    sc.resize(n); sh.resize(n);
    for (int i = 0; i < n; ++i) {
        double s = (double)g[i] / std::sqrt((double)v[i] + (double)BN_EPS);
        sc[i] = (float)s;
        sh[i] = (float)((double)b[i] - (double)m[i] * s);
    }
}

static int chunk_cap_default() {
    const char* e = getenv("IVOSW_CHUNK");
    int v = e ? atoi(e) : 128;
    return v < 1 ? 1 : v;
}

// workspace for `cap` scoring units (fp32 NHWC validation path)
static int ensure_workspace(ivosw_ctx* c, int cap) {
    int rc;
    const size_t f = sizeof(float);
    if ((rc = ensure(c->boxes, (size_t)cap * 4 * f))) return rc;
    if (c->conv_mode == IVOSW_CONV_SIMT_FP32) {
        if ((rc = ensure(c->crop, (size_t)cap * ROI * ROI * 4 * f))) return rc;
    } else {
        // padded split-fp16 crop planes; the zero border is written once per (re)allocation and never touched again
        const size_t plane = (size_t)cap * CROP_PH * CROP_PW * 8;
        if (c->crop_hi.bytes < plane || c->crop_lo.bytes < plane) {
            if ((rc = ensure(c->crop_hi, plane))) return rc;
            if ((rc = ensure(c->crop_lo, plane))) return rc;
            IVOSW_CUDA(cudaMemset(c->crop_hi.p, 0, c->crop_hi.bytes));
            IVOSW_CUDA(cudaMemset(c->crop_lo.p, 0, c->crop_lo.bytes));
        }
    }
    if ((rc = ensure(c->c1, (size_t)cap * 128 * 128 * 64 * f))) return rc;
    if ((rc = ensure(c->pool, (size_t)cap * 64 * 64 * 64 * f))) return rc;
    const size_t big = (size_t)cap * 64 * 64 * 256 * f;   // largest block output (res2)
    if ((rc = ensure(c->actX, big))) return rc;
    if ((rc = ensure(c->actY, big))) return rc;
    if ((rc = ensure(c->actDS, big))) return rc;
    const size_t mid = (size_t)cap * 64 * 64 * 128 * f;   // largest conv1/conv2 output (res3.0.conv1)
    if ((rc = ensure(c->actT1, mid))) return rc;
    if ((rc = ensure(c->actT2, mid))) return rc;
    return IVOSW_OK;
}

static int keep_probe(ivosw_ctx* c, int which, const float* src, size_t n_floats, cudaStream_t s) {
    if (!c->probes_on) return IVOSW_OK;
    int rc;
    if ((rc = ensure(c->probe_buf[which], n_floats * sizeof(float)))) return rc;
    IVOSW_CUDA(cudaMemcpyAsync(c->probe_buf[which].p, src, n_floats * sizeof(float), cudaMemcpyDeviceToDevice, s));
    return IVOSW_OK;
}

// Scores `n_units` units described by `ua` (ua.u0 is advanced per chunk) into score_dev[0..n_units).
static int assess_units(ivosw_ctx* c, UnitAddr ua, int n_units, int H, int W, float* score_dev, float* boxes_dev,
                        cudaStream_t s) {
    if (!c->assess_loaded) { set_error("AssessNet weights not loaded"); return IVOSW_ERR_STATE; }
    int rc;
    const int cap = std::min(n_units, c->chunk_cap);
    if ((rc = ensure_workspace(c, cap))) return rc;
    const int u_first = ua.u0;
    for (int done = 0; done < n_units; done += cap) {
        const int B = std::min(cap, n_units - done);
        ua.u0 = u_first + done;
        int tk = stage_begin(c, 0, s);
        if ((rc = launch_bbox(c, ua, B, H, W, s))) return rc;
        const bool tc = c->conv_mode != IVOSW_CONV_SIMT_FP32;
        const int terms = c->conv_mode == IVOSW_CONV_TC_FP16X3 ? 3 : 1;
        if ((rc = launch_roi_sample(c, ua, B, H, W, boxes_dev ? boxes_dev + 4 * (size_t)done : (float*)c->boxes.p, tc, s)))
            return rc;
        stage_end(c, tk, s);
        if (c->probes_on) {
            if (tc) {
                if ((rc = ensure(c->probe_buf[0], (size_t)B * ROI * ROI * 4 * sizeof(float)))) return rc;
                if ((rc = launch_crop_merge(c, (float*)c->probe_buf[0].p, B, terms == 3, s))) return rc;
            } else if ((rc = keep_probe(c, 0, (const float*)c->crop.p, (size_t)B * ROI * ROI * 4, s))) return rc;
        }
        SplitAct xs = split_view(c->c1);     // tensor-core path: pooled stem output as split-fp16 planes
        tk = stage_begin(c, 1, s);
        if (tc) {
            if ((rc = launch_stem_tc(c, B, xs, terms, s))) return rc;
        } else {
            if ((rc = launch_stem(c, B, s))) return rc;
        }
        stage_end(c, tk, s);
        if (c->probes_on) {
            const size_t n = (size_t)B * 64 * 64 * 64;
            if (tc) {
                if ((rc = ensure(c->probe_buf[1], n * sizeof(float)))) return rc;
                if ((rc = launch_merge(c, xs, (float*)c->probe_buf[1].p, (long long)n, terms == 3, s))) return rc;
            } else if ((rc = keep_probe(c, 1, (const float*)c->pool.p, n, s))) return rc;
        }
        tk = stage_begin(c, 2, s);
        const long long launches_before = c->launches;
        const float* x = nullptr;      // fp32 NHWC r5 handed to the pooling/FC kernel (CUDA-core path)
        SplitAct xs_final{nullptr, nullptr};   // split-fp16 r5 (tensor-core path)
        if (c->conv_mode == IVOSW_CONV_SIMT_FP32) {
            x = (const float*)c->pool.p;
            float* outs[2] = {(float*)c->actX.p, (float*)c->actY.p};
            int flip = 0, stage_probe = 2;
            for (size_t li = 0; li < c->layers.size(); ++li) {
                const ConvLayer& L = c->layers[li];
                if (L.first_of_block) {
                    if ((rc = launch_conv_simt(c, L, x, nullptr, (float*)c->actT1.p, B, s))) return rc;
                } else if (L.k == 3) {
                    if ((rc = launch_conv_simt(c, L, (const float*)c->actT1.p, nullptr, (float*)c->actT2.p, B, s))) return rc;
                } else if (L.is_downsample) {
                    if ((rc = launch_conv_simt(c, L, x, nullptr, (float*)c->actDS.p, B, s))) return rc;
                } else {  // conv3 + residual + relu -> block output
                    const float* res = L.residual == 2 ? (const float*)c->actDS.p : x;
                    float* y = outs[flip];
                    if ((rc = launch_conv_simt(c, L, (const float*)c->actT2.p, res, y, B, s))) return rc;
                    x = y;
                    flip ^= 1;
                    // a stage ends where the next bottleneck opens with a downsample branch
                    const bool stage_end_ = (li + 1 == c->layers.size()) ||
                                            (li + 3 < c->layers.size() && c->layers[li + 3].is_downsample);
                    if (stage_end_) {
                        if ((rc = keep_probe(c, stage_probe, x, (size_t)B * L.out_hw * L.out_hw * L.cout, s))) return rc;
                        ++stage_probe;
                    }
                }
            }
        } else {
            if (c->stack_on) {
                // one persistent launch for all 52 layers (conv_stack.cu)
                SplitAct stage_out[4];
                if ((rc = launch_conv_stack(c, xs, B, terms, s, &xs, c->probes_on ? stage_out : nullptr))) return rc;
                if (c->probes_on) {
                    static const int C_[4] = {256, 512, 1024, 2048}, HW_[4] = {64, 32, 16, 8};
                    for (int k = 0; k < 4; ++k) {
                        const size_t n = (size_t)B * HW_[k] * HW_[k] * C_[k];
                        if ((rc = ensure(c->probe_buf[2 + k], n * sizeof(float)))) return rc;
                        if ((rc = launch_merge(c, stage_out[k], (float*)c->probe_buf[2 + k].p, (long long)n, terms == 3, s)))
                            return rc;
                    }
                }
            } else {
            // tensor-core path: activations live as split-fp16 planes inside the same workspace buffers
                const SplitAct t1 = split_view(c->actT1), t2 = split_view(c->actT2), ds = split_view(c->actDS);
                const SplitAct outs[2] = {split_view(c->actX), split_view(c->actY)};
                int flip = 0, stage_probe = 2, fused_idx = 0;
                for (size_t li = 0; li < c->layers.size(); ++li) {
                    const ConvLayer& L = c->layers[li];
                    if (L.first_of_block) {
                        if ((rc = launch_conv_tc(c, L, xs, nullptr, t1, B, terms, s))) return rc;
                    } else if (L.k == 3) {
                        if ((rc = launch_conv_tc(c, L, t1, nullptr, t2, B, terms, s))) return rc;
                    } else if (L.is_downsample) {
                        // fused into the conv3 that follows (one GEMM over [t2 ; x]) unless IVOSW_FUSE_DS=0
                        if (!c->fuse_ds && (rc = launch_conv_tc(c, L, xs, nullptr, ds, B, terms, s))) return rc;
                    } else {
                        const SplitAct y = outs[flip];
                        if (L.residual == 2 && c->fuse_ds) {
                            if ((rc = launch_conv_tc(c, L, t2, nullptr, y, B, terms, s, &c->fused_tail[fused_idx], &xs))) return rc;
                        } else {
                            const SplitAct* res = L.residual == 2 ? &ds : &xs;
                            if ((rc = launch_conv_tc(c, L, t2, res, y, B, terms, s))) return rc;
                        }
                        if (L.residual == 2) ++fused_idx;
                        xs = y;
                        flip ^= 1;
                        const bool stage_end_ = (li + 1 == c->layers.size()) ||
                                                (li + 3 < c->layers.size() && c->layers[li + 3].is_downsample);
                        if (stage_end_) {
                            if (c->probes_on) {
                                const size_t n = (size_t)B * L.out_hw * L.out_hw * L.cout;
                                if ((rc = ensure(c->probe_buf[stage_probe], n * sizeof(float)))) return rc;
                                if ((rc = launch_merge(c, xs, (float*)c->probe_buf[stage_probe].p, (long long)n, terms == 3, s)))
                                    return rc;
                            }
                            ++stage_probe;
                        }
                    }
                }
            }
            xs_final = xs;
        }
        stage_end(c, tk, s);
        if (tk >= 0) c->conv_launches_timed += c->launches - launches_before;
        tk = stage_begin(c, 3, s);
        if (tc) {
            if ((rc = launch_gap_fc_split(c, xs_final, terms == 3, B, score_dev + done, s))) return rc;
        } else if ((rc = launch_gap_fc(c, x, B, score_dev + done, s))) return rc;
        stage_end(c, tk, s);
        c->last_chunk_b = B;
    }
    return IVOSW_OK;
}

int ensure_pinned(ivosw_ctx* c, size_t bytes) {
    if (c->pinned_small_bytes >= bytes) return IVOSW_OK;
    if (c->pinned_small) cudaFreeHost(c->pinned_small);
    c->pinned_small = nullptr; c->pinned_small_bytes = 0;
    ++g_alloc_epoch;
    IVOSW_CUDA(cudaMallocHost(&c->pinned_small, bytes));
    c->pinned_small_bytes = bytes;
    return IVOSW_OK;
}

// field by field: the struct has tail padding, which brace-initialised stack keys leave indeterminate
static bool same_key(const ivosw_ctx::GraphKey& a, const ivosw_ctx::GraphKey& b) {
    return a.p0 == b.p0 && a.p1 == b.p1 && a.p2 == b.p2 && a.T == b.T && a.O == b.O && a.H == b.H && a.W == b.W &&
           a.tb == b.tb && a.te == b.te && a.mode == b.mode && a.kind == b.kind && a.flags == b.flags;
}

static void drop_graph(ivosw_ctx* c, ivosw_ctx::GraphEntry& e) {
    if (c->last_graph == &e) drain_graph_events(c);
    if (e.exec) cudaGraphExecDestroy(e.exec);
    e.exec = nullptr;
    for (auto& ev : e.evts) { c->evt_pool.push_back(ev.a); c->evt_pool.push_back(ev.b); }
    e.evts.clear();
    e.seen = 0;
}

// Runs `fn(stream)` (which only ENQUEUES work) either eagerly or — from the second identical call on —
// as a replay of a CUDA graph captured from it.  First call: eager (it also performs every
// allocation); second call: stream capture + instantiate; afterwards: one cudaGraphLaunch per call.
// Graphs are rebuilt whenever a workspace buffer has moved since the capture.
template <class F>
static int run_graphed(ivosw_ctx* c, ivosw_ctx::GraphKey key, cudaStream_t s, F&& fn) {
    if (!c->graphs_on || c->probes_on) return fn(s);
    key.flags |= c->timing_on ? 0x100 : 0;
    ivosw_ctx::GraphEntry* e = nullptr;
    for (auto& g : c->graphs) if (same_key(g.key, key)) { e = &g; break; }
    if (e && e->epoch != g_alloc_epoch) { drop_graph(c, *e); }
    if (!e) {
        if (c->graphs.size() >= 64) { for (auto& g : c->graphs) drop_graph(c, g); c->graphs.clear(); }
        c->graphs.reserve(64);     // entries are referenced by pointer (last_graph): never reallocate
        c->graphs.emplace_back();
        e = &c->graphs.back();
        e->key = key;
    }
    if (e->seen == 0) {
        const int rc = fn(s);
        if (rc == IVOSW_OK) { e->seen = 1; e->epoch = g_alloc_epoch; }    // a failed eager run is not a template for capture
        return rc;
    }
    if (!e->exec) {
        if (!c->graph_stream) {
            IVOSW_CUDA(cudaStreamCreate(&c->graph_stream));
            IVOSW_CUDA(cudaEventCreateWithFlags(&c->g_ev1, cudaEventDisableTiming));
            IVOSW_CUDA(cudaEventCreateWithFlags(&c->g_ev2, cudaEventDisableTiming));
        }
        const unsigned long long epoch0 = g_alloc_epoch;
        const long long l0 = c->launches, cl0 = c->conv_launches_timed;
        IVOSW_CUDA(cudaStreamBeginCapture(c->graph_stream, cudaStreamCaptureModeThreadLocal));
        c->capturing = true; c->capture_evts = &e->evts;
        const int rc = fn(c->graph_stream);
        c->capturing = false; c->capture_evts = nullptr;
        cudaGraph_t g = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(c->graph_stream, &g);
        e->launches = c->launches - l0; e->conv_launches = c->conv_launches_timed - cl0;
        c->launches = l0; c->conv_launches_timed = cl0;
        if (rc != IVOSW_OK || ce != cudaSuccess || g == nullptr || epoch0 != g_alloc_epoch) {
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            drop_graph(c, *e);
            return fn(s);           // could not capture (a buffer had to grow, a plan had to be rebuilt): run eagerly
        }
        const cudaError_t ie = cudaGraphInstantiate(&e->exec, g, 0);
        cudaGraphDestroy(g);
        if (ie != cudaSuccess) { e->exec = nullptr; cudaGetLastError(); drop_graph(c, *e); return fn(s); }
        e->epoch = g_alloc_epoch;
    }
    if (c->last_graph) drain_graph_events(c);       // never overwrite unread timing events
    IVOSW_CUDA(cudaEventRecord(c->g_ev1, s));
    IVOSW_CUDA(cudaStreamWaitEvent(c->graph_stream, c->g_ev1, 0));
    IVOSW_CUDA(cudaGraphLaunch(e->exec, c->graph_stream));
    IVOSW_CUDA(cudaEventRecord(c->g_ev2, c->graph_stream));
    IVOSW_CUDA(cudaStreamWaitEvent(s, c->g_ev2, 0));
    c->launches += e->launches;
    if (c->timing_on) { c->conv_launches_timed += e->conv_launches; c->last_graph = e; }
    return IVOSW_OK;
}

}  // namespace ivosw

using namespace ivosw;

extern "C" {

int ivosw_abi_version(void) { return IVOSW_ABI_VERSION; }
const char* ivosw_last_error(void) { return g_err.c_str(); }

int ivosw_create(int device, int conv_mode, ivosw_ctx** out) {
    IVOSW_REQUIRE(out != nullptr, "out is NULL");
    IVOSW_REQUIRE(conv_mode >= 0 && conv_mode <= 2, "conv_mode");
    int n = 0;
    IVOSW_CUDA(cudaGetDeviceCount(&n));
    IVOSW_REQUIRE(device >= 0 && device < n, "device index");
    IVOSW_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    IVOSW_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error(std::string("ivosw_b200 is built for sm_100a only; device is ") + prop.name);
        return IVOSW_ERR_STATE;
    }
    ivosw_ctx* c = new ivosw_ctx();
    c->device = device;
    c->conv_mode = conv_mode;
    c->sm_count = prop.multiProcessorCount;
    c->chunk_cap = chunk_cap_default();
    { const char* g = getenv("IVOSW_GRAPHS"); c->graphs_on = !(g && atoi(g) == 0); }
    { const char* g = getenv("IVOSW_FUSE_DS"); c->fuse_ds = !(g && atoi(g) == 0); }
    { const char* g = getenv("IVOSW_STACK"); c->stack_on = g && atoi(g) != 0; }   // 1: conv_stack.cu (one persistent launch) instead of one launch per layer
    c->layers = make_resnet50_layers();
    if (cudaMalloc(&c->sat_count, sizeof(unsigned long long)) != cudaSuccess ||
        cudaMemset(c->sat_count, 0, sizeof(unsigned long long)) != cudaSuccess) {
        delete c;
        return cuda_fail(cudaGetLastError(), "saturation counter", __FILE__, __LINE__);
    }
    *out = c;
    return IVOSW_OK;
}

void ivosw_destroy(ivosw_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    if (c->brain_params) cudaFree(c->brain_params);
    if (c->brain_whh_t) cudaFree(c->brain_whh_t);
    if (c->brain_d1t) cudaFree(c->brain_d1t);
    if (c->brain_done_count) cudaFree(c->brain_done_count);
    if (c->target_params) cudaFree(c->target_params);
    if (c->target_whh_t) cudaFree(c->target_whh_t);
    if (c->target_d1t) cudaFree(c->target_d1t);
    if (c->adam_m) cudaFree(c->adam_m);
    if (c->adam_v) cudaFree(c->adam_v);
    release(c->dqn_ws);
    if (c->sat_count) cudaFree(c->sat_count);
    conv_stack_release(c);
    gather_release(c);
    train_release(c);
    manet_encoder_release(c);
    release(c->stack_arena);
    if (c->stem_w) cudaFree(c->stem_w);
    if (c->stem_scale) cudaFree(c->stem_scale);
    if (c->stem_shift) cudaFree(c->stem_shift);
    if (c->stem_wpack) cudaFree(c->stem_wpack);
    if (c->fc_w) cudaFree(c->fc_w);
    for (FusedTail& F : c->fused_tail) {
        if (F.w_hi) cudaFree(F.w_hi);
        if (F.w_lo) cudaFree(F.w_lo);
        if (F.scale) cudaFree(F.scale);
        if (F.shift) cudaFree(F.shift);
    }
    for (ConvLayer& L : c->layers) {
        if (L.w_f32) cudaFree(L.w_f32);
        if (L.scale) cudaFree(L.scale);
        if (L.shift) cudaFree(L.shift);
        if (L.w_hi) cudaFree(L.w_hi);
        if (L.w_lo) cudaFree(L.w_lo);
    }
    DeviceBuffer* bufs[] = {&c->brain_gi, &c->brain_h, &c->brain_state, &c->brain_q, &c->brain_arg, &c->bbox_min,
                            &c->bbox_max, &c->boxes, &c->crop, &c->crop_hi, &c->crop_lo, &c->c1, &c->pool, &c->actX, &c->actY, &c->actDS,
                            &c->actT1, &c->actT2, &c->scores, &c->scores_all, &c->mq, &c->stage_frames, &c->stage_probs, &c->band_min, &c->band_max,
                            &c->band_rows};
    for (DeviceBuffer* b : bufs) release(*b);
    for (DeviceBuffer& b : c->probe_buf) release(b);
    for (auto& g : c->graphs) drop_graph(c, g);
    c->graphs.clear();
    if (c->graph_stream) cudaStreamDestroy(c->graph_stream);
    if (c->g_ev1) cudaEventDestroy(c->g_ev1);
    if (c->g_ev2) cudaEventDestroy(c->g_ev2);
    drain_stage_events(c);
    for (cudaEvent_t e : c->evt_pool) cudaEventDestroy(e);
    if (c->pinned_small) cudaFreeHost(c->pinned_small);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->aux_stream) cudaStreamDestroy(c->aux_stream);
    if (c->pinned_rows) cudaFreeHost(c->pinned_rows);
    for (cudaEvent_t e : c->chunk_evts) cudaEventDestroy(e);
    delete c;
}

int ivosw_set_conv_mode(ivosw_ctx* c, int conv_mode) {
    IVOSW_REQUIRE(c != nullptr, "ctx");
    IVOSW_REQUIRE(conv_mode >= 0 && conv_mode <= 2, "conv_mode");
    c->conv_mode = conv_mode;
    return IVOSW_OK;
}

long long ivosw_launch_count(const ivosw_ctx* c) { return c ? c->launches : 0; }

long long ivosw_last_h2d_bytes(const ivosw_ctx* c) { return c ? c->last_h2d_bytes : 0; }

int ivosw_enable_probes(ivosw_ctx* c, int enable) {
    IVOSW_REQUIRE(c != nullptr, "ctx");
    c->probes_on = enable != 0;
    return IVOSW_OK;
}

// ------------------------------------------------------------------------------------------- Brain
int ivosw_brain_load(ivosw_ctx* c, const float* params_host, size_t n_floats) {
    IVOSW_REQUIRE(c && params_host, "null pointer");
    IVOSW_REQUIRE(n_floats == IVOSW_BRAIN_NUM_PARAMS, "Brain blob must hold 180993 floats");
    IVOSW_CUDA(cudaSetDevice(c->device));
    int rc;
    if ((rc = upload(&c->brain_params, params_host, n_floats))) return rc;
    if ((rc = brain_pack(c))) return rc;
    c->brain_loaded = true;
    // new policy weights from the host (a fresh Agent, load_state_dict): the Adam moments and step count belonged to the
    // previous weights.  Drop-in callers that only re-sync what the library itself produced do not come through here
    // (dropin/models/agent.py refreshes its stamp after update_agent).
    if (c->adam_m) {
        IVOSW_CUDA(cudaMemset(c->adam_m, 0, sizeof(float) * IVOSW_BRAIN_NUM_PARAMS));
        IVOSW_CUDA(cudaMemset(c->adam_v, 0, sizeof(float) * IVOSW_BRAIN_NUM_PARAMS));
    }
    c->adam_step = 0;
    return IVOSW_OK;
}

int ivosw_brain_forward(ivosw_ctx* c, const float* state_dev, int N, int T, float* q_dev, int* argmax_dev,
                        void* stream) {
    IVOSW_REQUIRE(c && state_dev && q_dev, "null pointer");
    IVOSW_REQUIRE(N >= 1 && T >= 1 && N <= 65535, "N, T");
    if (!c->brain_loaded) { set_error("Brain weights not loaded"); return IVOSW_ERR_STATE; }
    IVOSW_CUDA(cudaSetDevice(c->device));
    return launch_brain(c, state_dev, N, T, q_dev, argmax_dev, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------- DQN training step
int ivosw_dqn_load_target(ivosw_ctx* c, const float* params_host, size_t n_floats) {
    IVOSW_REQUIRE(c && params_host, "null pointer");
    IVOSW_REQUIRE(n_floats == IVOSW_BRAIN_NUM_PARAMS, "Brain blob must hold 180993 floats");
    IVOSW_CUDA(cudaSetDevice(c->device));
    int rc;
    if ((rc = upload(&c->target_params, params_host, n_floats))) return rc;
    if (!c->target_whh_t) IVOSW_CUDA(cudaMalloc(&c->target_whh_t, sizeof(float4) * 32 * 512));
    if (!c->target_d1t) IVOSW_CUDA(cudaMalloc(&c->target_d1t, sizeof(float) * 128 * 256));
    if ((rc = brain_pack_into(c, c->target_params, c->target_whh_t, c->target_d1t, nullptr))) return rc;
    IVOSW_CUDA(cudaDeviceSynchronize());
    c->target_loaded = true;
    return IVOSW_OK;
}

int ivosw_dqn_sync_target(ivosw_ctx* c, void* stream) {
    IVOSW_REQUIRE(c != nullptr, "ctx");
    if (!c->brain_loaded) { set_error("Brain weights not loaded"); return IVOSW_ERR_STATE; }
    IVOSW_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (!c->target_params) IVOSW_CUDA(cudaMalloc(&c->target_params, sizeof(float) * IVOSW_BRAIN_NUM_PARAMS));
    if (!c->target_whh_t) IVOSW_CUDA(cudaMalloc(&c->target_whh_t, sizeof(float4) * 32 * 512));
    if (!c->target_d1t) IVOSW_CUDA(cudaMalloc(&c->target_d1t, sizeof(float) * 128 * 256));
    IVOSW_CUDA(cudaMemcpyAsync(c->target_params, c->brain_params, sizeof(float) * IVOSW_BRAIN_NUM_PARAMS,
                               cudaMemcpyDeviceToDevice, s));
    int rc;
    if ((rc = brain_pack_into(c, c->target_params, c->target_whh_t, c->target_d1t, s))) return rc;
    c->target_loaded = true;
    return IVOSW_OK;
}

int ivosw_dqn_reset_optimizer(ivosw_ctx* c) {
    IVOSW_REQUIRE(c != nullptr, "ctx");
    IVOSW_CUDA(cudaSetDevice(c->device));
    if (c->adam_m) {
        IVOSW_CUDA(cudaMemset(c->adam_m, 0, sizeof(float) * IVOSW_BRAIN_NUM_PARAMS));
        IVOSW_CUDA(cudaMemset(c->adam_v, 0, sizeof(float) * IVOSW_BRAIN_NUM_PARAMS));
    }
    c->adam_step = 0;
    return IVOSW_OK;
}

int ivosw_dqn_update(ivosw_ctx* c, const float* state_dev, const float* new_state_dev, const int* action_dev,
                     const float* reward_step_dev, const float* reward_done_dev, int N, int T, float gamma, float lr,
                     float weight_decay, float* loss_host, float* grads_dev, int apply_update, void* stream) {
    IVOSW_REQUIRE(c && state_dev && new_state_dev && action_dev && reward_step_dev && reward_done_dev, "null pointer");
    IVOSW_REQUIRE(N >= 1 && N <= 65535 && T >= 1, "N, T");
    IVOSW_REQUIRE(apply_update || grads_dev, "grads_dev is required when the optimiser step is deferred");
    IVOSW_CUDA(cudaSetDevice(c->device));
    return dqn_update(c, state_dev, new_state_dev, action_dev, reward_step_dev, reward_done_dev, N, T, gamma, lr,
                      weight_decay, loss_host, grads_dev, apply_update != 0, (cudaStream_t)stream);
}

int ivosw_dqn_apply(ivosw_ctx* c, float* grads_dev, float lr, float weight_decay, void* stream) {
    IVOSW_REQUIRE(c && grads_dev, "null pointer");
    if (!c->brain_loaded) { set_error("Brain weights not loaded"); return IVOSW_ERR_STATE; }
    IVOSW_CUDA(cudaSetDevice(c->device));
    return dqn_apply(c, grads_dev, lr, weight_decay, (cudaStream_t)stream);
}

int ivosw_brain_get_params(ivosw_ctx* c, int which, float* out_dev, void* stream) {
    IVOSW_REQUIRE(c && out_dev, "null pointer");
    const float* src = which == 0 ? c->brain_params : c->target_params;
    if (!src) { set_error("requested Brain parameter set is not loaded"); return IVOSW_ERR_STATE; }
    IVOSW_CUDA(cudaSetDevice(c->device));
    IVOSW_CUDA(cudaMemcpyAsync(out_dev, src, sizeof(float) * IVOSW_BRAIN_NUM_PARAMS, cudaMemcpyDeviceToDevice,
                               (cudaStream_t)stream));
    return IVOSW_OK;
}

// --------------------------------------------------------------------------------------- AssessNet
size_t ivosw_assess_blob_floats(void) { return assess_blob_floats(); }

int ivosw_assess_load(ivosw_ctx* c, const float* blob, size_t n_floats) {
    IVOSW_REQUIRE(c && blob, "null pointer");
    IVOSW_REQUIRE(n_floats == assess_blob_floats(), "AssessNet blob length");
    IVOSW_CUDA(cudaSetDevice(c->device));
    int rc;
    const float* p = blob;
    for (int i = 0; i < 3; ++i) { c->mean[i] = p[i]; c->stdv[i] = p[3 + i]; }
    p += 6;
    {   // stem: [64][7][7][4] -> [(kh,kw,cin)][64]
        std::vector<float> wt((size_t)196 * 64);
        for (int co = 0; co < 64; ++co)
            for (int k = 0; k < 196; ++k) wt[(size_t)k * 64 + co] = p[(size_t)co * 196 + k];
        if ((rc = upload(&c->stem_w, wt.data(), wt.size()))) return rc;
        if ((rc = stem_tc_pack(c, p))) return rc;
        p += (size_t)64 * 196;
        std::vector<float> sc, sh;
        fold_bn(p, p + 64, p + 128, p + 192, 64, sc, sh);
        if ((rc = upload(&c->stem_scale, sc.data(), 64))) return rc;
        if ((rc = upload(&c->stem_shift, sh.data(), 64))) return rc;
        p += 256;
    }
    for (ConvLayer& L : c->layers) {
        const size_t nw = (size_t)L.cout * L.k * L.k * L.cin;
        if ((rc = upload(&L.w_f32, p, nw))) return rc;
        {   // split-fp16 copy for the tensor-core path: w ~= hi + lo / 2048
            std::vector<__half> hi(nw), lo(nw);
            for (size_t i = 0; i < nw; ++i) {
                __half h = __float2half_rn(p[i]);
                hi[i] = h;
                lo[i] = __float2half_rn((p[i] - __half2float(h)) * 2048.0f);
            }
            if (!L.w_hi) IVOSW_CUDA(cudaMalloc(&L.w_hi, nw * sizeof(__half)));
            if (!L.w_lo) IVOSW_CUDA(cudaMalloc(&L.w_lo, nw * sizeof(__half)));
            IVOSW_CUDA(cudaMemcpy(L.w_hi, hi.data(), nw * sizeof(__half), cudaMemcpyHostToDevice));
            IVOSW_CUDA(cudaMemcpy(L.w_lo, lo.data(), nw * sizeof(__half), cudaMemcpyHostToDevice));
        }
        p += nw;
        std::vector<float> sc, sh;
        fold_bn(p, p + L.cout, p + 2 * L.cout, p + 3 * L.cout, L.cout, sc, sh);
        if ((rc = upload(&L.scale, sc.data(), L.cout))) return rc;
        if ((rc = upload(&L.shift, sh.data(), L.cout))) return rc;
        p += 4 * (size_t)L.cout;
    }
    {   // conv3 + downsample of each stage's first bottleneck as one GEMM: [s3 W3 | sd Wd], shift b3 + bd (conv_tc.cu)
        const float* q = blob + 6 + (size_t)64 * 196 + 256;
        std::vector<const float*> wp(c->layers.size()), bnp(c->layers.size());
        for (size_t li = 0; li < c->layers.size(); ++li) {
            const ConvLayer& L = c->layers[li];
            wp[li] = q; q += (size_t)L.cout * L.k * L.k * L.cin;
            bnp[li] = q; q += 4 * (size_t)L.cout;
        }
        int fi = 0;
        for (size_t li = 0; li < c->layers.size(); ++li) {
            if (!c->layers[li].is_downsample) continue;
            const ConvLayer& D = c->layers[li];
            const ConvLayer& C3 = c->layers[li + 1];
            FusedTail& F = c->fused_tail[fi++];
            const int cout = C3.cout, k3 = C3.cin, kd = D.cin, kt = k3 + kd;
            std::vector<float> sc3, sh3, scd, shd;
            fold_bn(bnp[li + 1], bnp[li + 1] + cout, bnp[li + 1] + 2 * cout, bnp[li + 1] + 3 * cout, cout, sc3, sh3);
            fold_bn(bnp[li], bnp[li] + cout, bnp[li] + 2 * cout, bnp[li] + 3 * cout, cout, scd, shd);
            std::vector<__half> hi((size_t)cout * kt), lo((size_t)cout * kt);
            std::vector<float> ones(cout, 1.0f), shf(cout);
            for (int o = 0; o < cout; ++o) {
                shf[o] = (float)((double)sh3[o] + (double)shd[o]);
                for (int k = 0; k < kt; ++k) {
                    const double w = k < k3 ? (double)sc3[o] * (double)wp[li + 1][(size_t)o * k3 + k]
                                            : (double)scd[o] * (double)wp[li][(size_t)o * kd + (k - k3)];
                    const float wf = (float)w;
                    const __half h = __float2half_rn(wf);
                    hi[(size_t)o * kt + k] = h;
                    lo[(size_t)o * kt + k] = __float2half_rn((wf - __half2float(h)) * 2048.0f);
                }
            }
            if (!F.w_hi) IVOSW_CUDA(cudaMalloc(&F.w_hi, hi.size() * sizeof(__half)));
            if (!F.w_lo) IVOSW_CUDA(cudaMalloc(&F.w_lo, lo.size() * sizeof(__half)));
            IVOSW_CUDA(cudaMemcpy(F.w_hi, hi.data(), hi.size() * sizeof(__half), cudaMemcpyHostToDevice));
            IVOSW_CUDA(cudaMemcpy(F.w_lo, lo.data(), lo.size() * sizeof(__half), cudaMemcpyHostToDevice));
            if ((rc = upload(&F.scale, ones.data(), cout))) return rc;
            if ((rc = upload(&F.shift, shf.data(), cout))) return rc;
            F.k_total = kt; F.cin2 = kd; F.in_hw2 = D.in_hw; F.stride2 = D.stride;
        }
    }
    if ((rc = upload(&c->fc_w, p, 2048))) return rc;
    c->fc_b = p[2048];
    c->assess_loaded = true;
    ++c->assess_version;
    ++g_alloc_epoch;   // scalars (fc bias, mean, std) are baked into captured kernel arguments
    return IVOSW_OK;
}

int ivosw_assess_forward(ivosw_ctx* c, const float* frames_dev, long long frame_stride, const float* prob_dev,
                         long long prob_stride, int B, int H, int W, float* score_dev, float* boxes_dev,
                         void* stream) {
    IVOSW_REQUIRE(c && frames_dev && prob_dev && score_dev, "null pointer");
    IVOSW_REQUIRE(B >= 1 && H >= 2 && W >= 2, "B, H, W");
    IVOSW_REQUIRE((long long)H * W < (1ll << 30), "frame too large");
    IVOSW_CUDA(cudaSetDevice(c->device));
    UnitAddr ua{frames_dev, frame_stride, prob_dev, prob_stride, 0, B, 0};
    return assess_units(c, ua, B, H, W, score_dev, boxes_dev, (cudaStream_t)stream);
}

int ivosw_assess_probe(ivosw_ctx* c, int which, float* out_dev, size_t capacity_floats, int* dims4_out,
                       void* stream) {
    IVOSW_REQUIRE(c && out_dev, "null pointer");
    IVOSW_REQUIRE(which >= 0 && which < 6, "which");
    if (!c->probes_on || !c->probe_buf[which].p) { set_error("probes not enabled / no forward yet"); return IVOSW_ERR_STATE; }
    static const int C_[6] = {4, 64, 256, 512, 1024, 2048}, HW_[6] = {256, 64, 64, 32, 16, 8};
    const int B = c->last_chunk_b;
    const size_t n = (size_t)B * C_[which] * HW_[which] * HW_[which];
    IVOSW_REQUIRE(capacity_floats >= n, "probe output too small");
    if (dims4_out) { dims4_out[0] = B; dims4_out[1] = C_[which]; dims4_out[2] = HW_[which]; dims4_out[3] = HW_[which]; }
    return launch_nhwc_to_nchw(c, (const float*)c->probe_buf[which].p, out_dev, B, HW_[which] * HW_[which], C_[which],
                               (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------- round
// scoring half of a round for frames [t_begin, t_end): asynchronous, device outputs only
static int score_shard(ivosw_ctx* c, const float* frames_dev, const float* probs_dev, int T, int O, int H, int W,
                       int t_begin, int t_end, const double* ann_dev, double* mq_dev, float* state_dev,
                       float* scores_dev_out, int scores_pitch, cudaStream_t s) {
    const int Tl = t_end - t_begin;
    const long long HW = (long long)H * W;
    int rc;
    if ((rc = ensure(c->scores, sizeof(float) * (size_t)Tl * O))) return rc;
    UnitAddr ua{frames_dev + (long long)t_begin * 3 * HW, 3 * HW,
                probs_dev + ((long long)t_begin * (O + 1) + 1) * HW, (long long)(O + 1) * HW, HW, Tl, 0};
    if ((rc = assess_units(c, ua, Tl * O, H, W, (float*)c->scores.p, nullptr, s))) return rc;
    int tk = stage_begin(c, 3, s);
    if ((rc = launch_object_mean(c, (const float*)c->scores.p, Tl, O, ann_dev ? ann_dev + t_begin : nullptr, mq_dev,
                                 ann_dev ? state_dev : nullptr, s)))
        return rc;
    stage_end(c, tk, s);
    if (scores_dev_out)   // [O][Tl] rows into a destination whose rows are scores_pitch floats apart
        IVOSW_CUDA(cudaMemcpy2DAsync(scores_dev_out, sizeof(float) * (size_t)scores_pitch, c->scores.p,
                                     sizeof(float) * (size_t)Tl, sizeof(float) * (size_t)Tl, O,
                                     cudaMemcpyDeviceToDevice, s));
    return IVOSW_OK;
}

static int timed_brain(ivosw_ctx* c, int T, cudaStream_t s) {
    int tk = stage_begin(c, 4, s);
    int rc = launch_brain(c, (const float*)c->brain_state.p, 1, T, (float*)c->brain_q.p, (int*)c->brain_arg.p, s);
    stage_end(c, tk, s);
    return rc;
}

static int e2e_chunk_frames() {
    const char* e = getenv("IVOSW_E2E_CHUNK_FRAMES");
    int v = e ? atoi(e) : 16;
    return v < 1 ? 1 : v;
}

// Scores frames [t_begin, t_end) whose inputs live in HOST memory: the frames and the O foreground
// probability planes are uploaded into the context's staging buffers in chunks of a few frames on a
// second stream; each chunk's scoring waits only for its own copy, so PCIe transfer and compute
// overlap.  Probability channel 0 (background) is never read by the path and is not transferred.
// mq_dev receives the float64 per-frame means; c->scores_all the per-object scores ([O][Tl]) on request.
static int score_range_from_host(ivosw_ctx* c, const float* frames_host, const float* probs_host, int T, int O,
                                 int H, int W, int t_begin, int t_end, double* mq_dev, bool keep_scores,
                                 cudaStream_t s) {
    const int Tl = t_end - t_begin;
    const size_t HW = (size_t)H * W;
    int rc;
    // chunk schedule (frames per upload + scoring pass)
    const int FC = e2e_chunk_frames();
    std::vector<int> bounds;
    if (const char* sch = getenv("IVOSW_E2E_SCHEDULE")) {      // measurement: explicit chunk sizes "4,12,16,..." (last repeats)
        int pos = t_begin, last = FC;
        const char* p = sch;
        while (pos < t_end) {
            if (*p) { last = std::max(1, atoi(p)); while (*p && *p != ',') ++p; if (*p == ',') ++p; }
            bounds.push_back(pos); pos += std::min(last, t_end - pos);
        }
        bounds.push_back(t_end);
    } else {
        // FC/2 first (the first scoring pass starts as early as possible), then FC per chunk, whatever is left
        // (at most FC) last.  Measured on C2 (scripts/ab_e2e.sh): 8,16,16,16,8 frames = 10.6 ms against 11.6 ms for
        // 16,16,16,8,4,4 — per-pass efficiency matters more than a short tail, the round is compute-bound once
        // only the ROI row bands are sent.
        int pos = t_begin;
        const int first = std::max(1, FC / 2);
        bounds.push_back(pos); pos += std::min(first, t_end - pos);
        while (pos < t_end) { bounds.push_back(pos); pos += std::min(FC, t_end - pos); }
        bounds.push_back(t_end);
    }
    const int n_chunks = (int)bounds.size() - 1;
    if (!c->copy_stream) IVOSW_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    if (!c->aux_stream) IVOSW_CUDA(cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking));
    while ((int)c->chunk_evts.size() < 3 * n_chunks + 1) {
        cudaEvent_t e;
        IVOSW_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->chunk_evts.push_back(e);
    }
    cudaEvent_t* ev_p = c->chunk_evts.data();                 // probability planes of chunk ci on the device
    cudaEvent_t* ev_b = c->chunk_evts.data() + n_chunks;      // row ranges of chunk ci in pinned memory
    cudaEvent_t* ev_f = c->chunk_evts.data() + 2 * n_chunks;  // frame rows of chunk ci on the device
    cudaEvent_t ev_start = c->chunk_evts[3 * n_chunks];
    if ((rc = ensure(c->stage_frames, sizeof(float) * (size_t)T * 3 * HW))) return rc;
    if ((rc = ensure(c->stage_probs, sizeof(float) * (size_t)T * (O + 1) * HW))) return rc;
    if (keep_scores && (rc = ensure(c->scores_all, sizeof(float) * (size_t)Tl * O))) return rc;
    if ((rc = ensure(c->band_rows, sizeof(int4) * (size_t)T))) return rc;
    if (c->pinned_rows_n < (size_t)T) {
        if (c->pinned_rows) cudaFreeHost(c->pinned_rows);
        c->pinned_rows = nullptr; c->pinned_rows_n = 0;
        IVOSW_CUDA(cudaMallocHost((void**)&c->pinned_rows, sizeof(int4) * (size_t)T));
        c->pinned_rows_n = (size_t)T;
    }
    float* fs = (float*)c->stage_frames.p;
    float* ps = (float*)c->stage_probs.p;
    // Only the rows of a frame that an ROI can touch are sent (IVOSW_E2E_BANDS=0: whole frames).  Per chunk: probability
    // planes -> pre-pass on a side stream (bbox + the sampler's own row arithmetic) -> 8 bytes per frame back to the host
    // -> one strided copy per frame for its row band -> scoring pass.  The link idles for the ~50 us of the pre-pass
    // round trip per chunk; queueing the next chunk's planes ahead of these rows instead delays the scoring pass by a
    // whole plane transfer and measured slower (12.1 vs 11.6 ms).
    const bool bands = !(getenv("IVOSW_E2E_BANDS") && atoi(getenv("IVOSW_E2E_BANDS")) == 0);
    const bool cols = !(getenv("IVOSW_E2E_COLS") && atoi(getenv("IVOSW_E2E_COLS")) == 0);   // 0: whole rows of the band
    // the copy stream must not overtake work already queued on s that may still read the staging buffers
    IVOSW_CUDA(cudaEventRecord(ev_start, s));
    IVOSW_CUDA(cudaStreamWaitEvent(c->copy_stream, ev_start, 0));
    if (getenv("IVOSW_E2E_POISON") && atoi(getenv("IVOSW_E2E_POISON")) != 0)   // tests: unsent rows must never be read
        IVOSW_CUDA(cudaMemsetAsync(fs, 0xff, sizeof(float) * (size_t)T * 3 * HW, c->copy_stream));
    long long sent = 0;
    auto send_probs = [&](int ci) -> int {
        const int c0 = bounds[ci], c1 = bounds[ci + 1];
        IVOSW_CUDA(cudaMemcpy2DAsync(ps + ((size_t)c0 * (O + 1) + 1) * HW, sizeof(float) * (size_t)(O + 1) * HW,
                                     probs_host + ((size_t)c0 * (O + 1) + 1) * HW, sizeof(float) * (size_t)(O + 1) * HW,
                                     sizeof(float) * (size_t)O * HW, (size_t)(c1 - c0), cudaMemcpyHostToDevice,
                                     c->copy_stream));
        sent += (long long)sizeof(float) * O * HW * (c1 - c0);
        IVOSW_CUDA(cudaEventRecord(ev_p[ci], c->copy_stream));
        if (!bands) return IVOSW_OK;
        int r;
        IVOSW_CUDA(cudaStreamWaitEvent(c->aux_stream, ev_p[ci], 0));
        UnitAddr ua{fs + (long long)c0 * 3 * (long long)HW, 3 * (long long)HW,
                    ps + ((long long)c0 * (O + 1) + 1) * (long long)HW, (long long)(O + 1) * (long long)HW, (long long)HW,
                    c1 - c0, 0};
        if ((r = launch_bbox(c, ua, (c1 - c0) * O, H, W, c->aux_stream, &c->band_min, &c->band_max))) return r;
        if ((r = launch_roi_rows(c, (const int2*)c->band_min.p, (const int2*)c->band_max.p, c1 - c0, O, H, W,
                                 (int4*)c->band_rows.p + c0, c->aux_stream)))
            return r;
        IVOSW_CUDA(cudaMemcpyAsync(c->pinned_rows + c0, (int4*)c->band_rows.p + c0, sizeof(int4) * (size_t)(c1 - c0),
                                   cudaMemcpyDeviceToHost, c->aux_stream));
        IVOSW_CUDA(cudaEventRecord(ev_b[ci], c->aux_stream));
        return IVOSW_OK;
    };
    if (bands) {      // band_min / band_max are sized once, for the largest chunk, before anything is queued on them
        int big = 0;
        for (int ci = 0; ci < n_chunks; ++ci) big = std::max(big, bounds[ci + 1] - bounds[ci]);
        if ((rc = ensure(c->band_min, sizeof(int2) * (size_t)big * O))) return rc;
        if ((rc = ensure(c->band_max, sizeof(int2) * (size_t)big * O))) return rc;
    }
    if ((rc = send_probs(0))) return rc;
    for (int ci = 0; ci < n_chunks; ++ci) {
        const int c0 = bounds[ci], c1 = bounds[ci + 1];
        if (bands) {
            IVOSW_CUDA(cudaEventSynchronize(ev_b[ci]));
            for (int t = c0; t < c1; ++t) {
                const int4 r = c->pinned_rows[t];           // first / last row, first / last column an ROI of this frame can touch
                if (r.x > r.y) continue;
                const int nrows = r.y - r.x + 1, ncols = r.w - r.z + 1;
                if (cols && ncols * 8 < W * 7) {
                    // the rectangle only: one strided copy per colour plane (rows of ncols floats, 128-byte aligned starts)
                    for (int ch = 0; ch < 3; ++ch) {
                        const size_t off = ((size_t)t * 3 + ch) * HW + (size_t)r.x * W + r.z;
                        IVOSW_CUDA(cudaMemcpy2DAsync(fs + off, sizeof(float) * W, frames_host + off, sizeof(float) * W,
                                                     sizeof(float) * (size_t)ncols, (size_t)nrows, cudaMemcpyHostToDevice,
                                                     c->copy_stream));
                    }
                    sent += 3ll * (long long)sizeof(float) * ncols * nrows;
                } else {
                    const size_t off = (size_t)t * 3 * HW + (size_t)r.x * W;
                    const size_t bytes = sizeof(float) * (size_t)nrows * W;
                    IVOSW_CUDA(cudaMemcpy2DAsync(fs + off, sizeof(float) * HW, frames_host + off, sizeof(float) * HW, bytes, 3,
                                                 cudaMemcpyHostToDevice, c->copy_stream));
                    sent += 3 * (long long)bytes;
                }
            }
        } else {
            IVOSW_CUDA(cudaMemcpyAsync(fs + (size_t)c0 * 3 * HW, frames_host + (size_t)c0 * 3 * HW,
                                       sizeof(float) * (size_t)(c1 - c0) * 3 * HW, cudaMemcpyHostToDevice, c->copy_stream));
            sent += (long long)sizeof(float) * 3 * HW * (c1 - c0);
            if (ci + 1 < n_chunks && (rc = send_probs(ci + 1))) return rc;
        }
        IVOSW_CUDA(cudaEventRecord(ev_f[ci], c->copy_stream));
        if (bands && ci + 1 < n_chunks && (rc = send_probs(ci + 1))) return rc;       // next planes right behind these rows
        IVOSW_CUDA(cudaStreamWaitEvent(s, ev_f[ci], 0));
        // the scoring pass of a chunk is an enqueue-only sequence over stable staging pointers: CUDA-graph replay
        // (the ~56 launches and ~300 tensor-map encodes of a pass otherwise cost more host time than the pass runs)
        auto enqueue = [&](cudaStream_t st) -> int {
            return score_shard(c, fs, ps, T, O, H, W, c0, c1, nullptr, mq_dev + (c0 - t_begin), nullptr,
                               keep_scores ? (float*)c->scores_all.p + (c0 - t_begin) : nullptr, Tl, st);
        };
        if (c->timing_on) {
            if ((rc = enqueue(s))) return rc;
        } else {
            if ((rc = ensure(c->scores, sizeof(float) * (size_t)(c1 - c0) * O))) return rc;
            ivosw_ctx::GraphKey key{fs, ps, mq_dev + (c0 - t_begin), T, O, H, W, c0, c1, c->conv_mode, 5, keep_scores ? 1 : 0};
            if ((rc = run_graphed(c, key, s, enqueue))) return rc;
        }
    }
    c->last_h2d_bytes = sent;
    return IVOSW_OK;
}

// One round over frames [t_begin, t_end).  With frames_host / probs_host set, the inputs are uploaded
// into the context's staging buffers in frame chunks on a second stream, each chunk's scoring
// waiting only for its own copy, so PCIe transfer and compute overlap; probability channel 0
// (background) is never read by the path and is not transferred.
static int round_core(ivosw_ctx* c, const float* frames_dev, const float* probs_dev, const float* frames_host,
                      const float* probs_host, int T, int O, int H, int W, int t_begin, int t_end,
                      const double* ann_host, double* mq_host, float* scores_host, float* q_host, int* next_frame,
                      cudaStream_t s) {
    const int Tl = t_end - t_begin;
    int rc;
    // mq buffer: [Tl doubles mq][T doubles ann]
    if ((rc = ensure(c->mq, sizeof(double) * ((size_t)Tl + T)))) return rc;
    if ((rc = ensure(c->brain_state, sizeof(float) * 2 * (size_t)T))) return rc;
    if ((rc = ensure(c->brain_q, sizeof(float) * (size_t)T))) return rc;
    if ((rc = ensure(c->brain_arg, sizeof(int)))) return rc;
    const size_t pin_bytes = sizeof(double) * (size_t)(Tl + T) + sizeof(float) * ((size_t)Tl * O + T) + 64;
    if ((rc = ensure_pinned(c, pin_bytes))) return rc;
    double* pin_mq = (double*)c->pinned_small;
    double* pin_ann = pin_mq + Tl;
    float* pin_scores = (float*)(pin_ann + T);
    float* pin_q = pin_scores + (size_t)Tl * O;
    int* pin_arg = (int*)(pin_q + T);

    double* mq_dev = (double*)c->mq.p;
    double* ann_dev = mq_dev + Tl;
    const bool full = (t_begin == 0 && t_end == T) && (q_host || next_frame);
    if (full && !c->brain_loaded) { set_error("Brain weights not loaded"); return IVOSW_ERR_STATE; }
    memcpy(pin_ann, ann_host, sizeof(double) * T);
    if (!frames_host) {
        // device-resident inputs: the whole round is one enqueue-only sequence -> CUDA-graph replay
        if ((rc = ensure(c->scores, sizeof(float) * (size_t)Tl * O))) return rc;
        auto enqueue = [&](cudaStream_t st) -> int {
            int r;
            IVOSW_CUDA(cudaMemcpyAsync(ann_dev, pin_ann, sizeof(double) * T, cudaMemcpyHostToDevice, st));
            if ((r = score_shard(c, frames_dev, probs_dev, T, O, H, W, t_begin, t_end, ann_dev, mq_dev,
                                 full ? (float*)c->brain_state.p : nullptr, nullptr, 0, st)))
                return r;
            if (full) {
                if ((r = timed_brain(c, T, st))) return r;
                IVOSW_CUDA(cudaMemcpyAsync(pin_q, c->brain_q.p, sizeof(float) * T, cudaMemcpyDeviceToHost, st));
                IVOSW_CUDA(cudaMemcpyAsync(pin_arg, c->brain_arg.p, sizeof(int), cudaMemcpyDeviceToHost, st));
            }
            IVOSW_CUDA(cudaMemcpyAsync(pin_mq, mq_dev, sizeof(double) * Tl, cudaMemcpyDeviceToHost, st));
            if (scores_host)
                IVOSW_CUDA(cudaMemcpyAsync(pin_scores, c->scores.p, sizeof(float) * (size_t)Tl * O,
                                           cudaMemcpyDeviceToHost, st));
            return IVOSW_OK;
        };
        ivosw_ctx::GraphKey key{frames_dev, probs_dev, nullptr, T, O, H, W, t_begin, t_end, c->conv_mode, 1,
                                (full ? 1 : 0) | (scores_host ? 2 : 0)};
        if ((rc = run_graphed(c, key, s, enqueue))) return rc;
    } else {
        IVOSW_CUDA(cudaMemcpyAsync(ann_dev, pin_ann, sizeof(double) * T, cudaMemcpyHostToDevice, s));
        if ((rc = score_range_from_host(c, frames_host, probs_host, T, O, H, W, t_begin, t_end, mq_dev,
                                        scores_host != nullptr, s)))
            return rc;
        if (full && (rc = launch_pack_state(c, mq_dev, ann_dev, T, (float*)c->brain_state.p, s))) return rc;
        if (full) {
            if ((rc = timed_brain(c, T, s))) return rc;
            IVOSW_CUDA(cudaMemcpyAsync(pin_q, c->brain_q.p, sizeof(float) * T, cudaMemcpyDeviceToHost, s));
            IVOSW_CUDA(cudaMemcpyAsync(pin_arg, c->brain_arg.p, sizeof(int), cudaMemcpyDeviceToHost, s));
        }
        IVOSW_CUDA(cudaMemcpyAsync(pin_mq, mq_dev, sizeof(double) * Tl, cudaMemcpyDeviceToHost, s));
        if (scores_host)
            IVOSW_CUDA(cudaMemcpyAsync(pin_scores, c->scores_all.p, sizeof(float) * (size_t)Tl * O,
                                       cudaMemcpyDeviceToHost, s));
    }
    IVOSW_CUDA(cudaStreamSynchronize(s));
    drain_graph_events(c);
    memcpy(mq_host, pin_mq, sizeof(double) * Tl);
    if (scores_host)   // device layout is [O][Tl]; the reference's mask_quality_pred is [Tl][O]
        for (int t = 0; t < Tl; ++t)
            for (int o = 0; o < O; ++o) scores_host[(size_t)t * O + o] = pin_scores[(size_t)o * Tl + t];
    if (full && q_host) memcpy(q_host, pin_q, sizeof(float) * T);
    if (full && next_frame) *next_frame = *pin_arg;
    return IVOSW_OK;
}

int ivosw_round_device(ivosw_ctx* c, const float* frames_dev, const float* probs_dev, int T, int O, int H, int W,
                       int t_begin, int t_end, const double* ann_host, double* mq_host, float* scores_host,
                       float* q_host, int* next_frame, void* stream) {
    IVOSW_REQUIRE(c && frames_dev && probs_dev && ann_host && mq_host, "null pointer");
    IVOSW_REQUIRE(T >= 1 && O >= 1 && H >= 2 && W >= 2, "T, O, H, W");
    IVOSW_REQUIRE(0 <= t_begin && t_begin < t_end && t_end <= T, "frame range");
    IVOSW_CUDA(cudaSetDevice(c->device));
    return round_core(c, frames_dev, probs_dev, nullptr, nullptr, T, O, H, W, t_begin, t_end, ann_host, mq_host,
                      scores_host, q_host, next_frame, (cudaStream_t)stream);
}

int ivosw_round_host(ivosw_ctx* c, const float* frames_host, const float* probs_host, int T, int O, int H, int W,
                     const double* ann_host, double* mq_host, float* scores_host, float* q_host, int* next_frame,
                     void* stream) {
    IVOSW_REQUIRE(c && frames_host && probs_host && ann_host && mq_host, "null pointer");
    IVOSW_REQUIRE(T >= 1 && O >= 1 && H >= 2 && W >= 2, "T, O, H, W");
    IVOSW_CUDA(cudaSetDevice(c->device));
    return round_core(c, nullptr, nullptr, frames_host, probs_host, T, O, H, W, 0, T, ann_host, mq_host, scores_host,
                      q_host, next_frame, (cudaStream_t)stream);
}

int ivosw_agent_action(ivosw_ctx* c, const double* mq_host, const double* ann_host, int T, float* q_host,
                       int* next_frame, void* stream) {
    IVOSW_REQUIRE(c && mq_host && ann_host, "null pointer");
    IVOSW_REQUIRE(T >= 1, "T");
    if (!c->brain_loaded) { set_error("Brain weights not loaded"); return IVOSW_ERR_STATE; }
    IVOSW_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    if ((rc = ensure(c->brain_state, sizeof(float) * 2 * (size_t)T))) return rc;
    if ((rc = ensure(c->brain_q, sizeof(float) * (size_t)T))) return rc;
    if ((rc = ensure(c->brain_arg, sizeof(int)))) return rc;
    if ((rc = ensure_pinned(c, sizeof(float) * 3 * (size_t)T + 64))) return rc;
    float* pin_state = (float*)c->pinned_small;
    float* pin_q = pin_state + 2 * (size_t)T;
    int* pin_arg = (int*)(pin_q + T);
    for (int t = 0; t < T; ++t) {   // torch.Tensor(state[None]) : float64 -> float32 (agent.py:176)
        pin_state[2 * t] = (float)mq_host[t];
        pin_state[2 * t + 1] = (float)ann_host[t];
    }
    IVOSW_CUDA(cudaMemcpyAsync(c->brain_state.p, pin_state, sizeof(float) * 2 * T, cudaMemcpyHostToDevice, s));
    if ((rc = launch_brain(c, (const float*)c->brain_state.p, 1, T, (float*)c->brain_q.p, (int*)c->brain_arg.p, s)))
        return rc;
    IVOSW_CUDA(cudaMemcpyAsync(pin_q, c->brain_q.p, sizeof(float) * T, cudaMemcpyDeviceToHost, s));
    IVOSW_CUDA(cudaMemcpyAsync(pin_arg, c->brain_arg.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    IVOSW_CUDA(cudaStreamSynchronize(s));
    if (q_host) memcpy(q_host, pin_q, sizeof(float) * T);
    if (next_frame) *next_frame = *pin_arg;
    return IVOSW_OK;
}

int ivosw_score_shard(ivosw_ctx* c, const float* frames_dev, const float* probs_dev, int T, int O, int H, int W,
                      int t_begin, int t_end, double* mq_dev, float* scores_dev, void* stream) {
    IVOSW_REQUIRE(c && frames_dev && probs_dev && mq_dev, "null pointer");
    IVOSW_REQUIRE(T >= 1 && O >= 1 && H >= 2 && W >= 2, "T, O, H, W");
    IVOSW_REQUIRE(0 <= t_begin && t_begin < t_end && t_end <= T, "frame range");
    IVOSW_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    auto enqueue = [&](cudaStream_t st) -> int {
        return score_shard(c, frames_dev, probs_dev, T, O, H, W, t_begin, t_end, nullptr, mq_dev, nullptr, scores_dev,
                           t_end - t_begin, st);
    };
    // (no host synchronisation here, so per-replay timing events could not be read back: eager when timing)
    if (c->timing_on || scores_dev) return enqueue(s);
    int rc;
    if ((rc = ensure(c->scores, sizeof(float) * (size_t)(t_end - t_begin) * O))) return rc;
    ivosw_ctx::GraphKey key{frames_dev, probs_dev, mq_dev, T, O, H, W, t_begin, t_end, c->conv_mode, 2, 0};
    return run_graphed(c, key, s, enqueue);
}

int ivosw_score_shard_host(ivosw_ctx* c, const float* frames_host, const float* probs_host, int T, int O, int H, int W,
                           int t_begin, int t_end, double* mq_dev, void* stream) {
    IVOSW_REQUIRE(c && frames_host && probs_host && mq_dev, "null pointer");
    IVOSW_REQUIRE(T >= 1 && O >= 1 && H >= 2 && W >= 2, "T, O, H, W");
    IVOSW_REQUIRE(0 <= t_begin && t_begin < t_end && t_end <= T, "frame range");
    IVOSW_CUDA(cudaSetDevice(c->device));
    return score_range_from_host(c, frames_host, probs_host, T, O, H, W, t_begin, t_end, mq_dev, false,
                                 (cudaStream_t)stream);
}

int ivosw_agent_action_dev(ivosw_ctx* c, const double* mq_dev, const double* ann_host, int T, float* q_host,
                           int* next_frame, void* stream) {
    IVOSW_REQUIRE(c && mq_dev && ann_host, "null pointer");
    IVOSW_REQUIRE(T >= 1, "T");
    if (!c->brain_loaded) { set_error("Brain weights not loaded"); return IVOSW_ERR_STATE; }
    IVOSW_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    if ((rc = ensure(c->mq, sizeof(double) * 2 * (size_t)T))) return rc;
    if ((rc = ensure(c->brain_state, sizeof(float) * 2 * (size_t)T))) return rc;
    if ((rc = ensure(c->brain_q, sizeof(float) * (size_t)T))) return rc;
    if ((rc = ensure(c->brain_arg, sizeof(int)))) return rc;
    if ((rc = ensure_pinned(c, sizeof(double) * T + sizeof(float) * T + 64))) return rc;
    double* pin_ann = (double*)c->pinned_small;
    float* pin_q = (float*)(pin_ann + T);
    int* pin_arg = (int*)(pin_q + T);
    double* ann_dev = (double*)c->mq.p + T;   // second half of the mq buffer (first half may be mq_dev itself)
    memcpy(pin_ann, ann_host, sizeof(double) * T);
    if ((rc = ensure(c->brain_gi, sizeof(float) * (size_t)T * 512))) return rc;
    if ((rc = ensure(c->brain_h, sizeof(float) * (size_t)2 * T * 128))) return rc;
    auto enqueue = [&](cudaStream_t st) -> int {
        int r;
        IVOSW_CUDA(cudaMemcpyAsync(ann_dev, pin_ann, sizeof(double) * T, cudaMemcpyHostToDevice, st));
        if ((r = launch_pack_state(c, mq_dev, ann_dev, T, (float*)c->brain_state.p, st))) return r;
        if ((r = timed_brain(c, T, st))) return r;
        IVOSW_CUDA(cudaMemcpyAsync(pin_q, c->brain_q.p, sizeof(float) * T, cudaMemcpyDeviceToHost, st));
        IVOSW_CUDA(cudaMemcpyAsync(pin_arg, c->brain_arg.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        return IVOSW_OK;
    };
    ivosw_ctx::GraphKey key{mq_dev, nullptr, nullptr, T, 0, 0, 0, 0, 0, 0, 3, 0};
    if ((rc = run_graphed(c, key, s, enqueue))) return rc;
    IVOSW_CUDA(cudaStreamSynchronize(s));
    drain_graph_events(c);
    if (q_host) memcpy(q_host, pin_q, sizeof(float) * T);
    if (next_frame) *next_frame = *pin_arg;
    return IVOSW_OK;
}

int ivosw_conv_saturation_count(ivosw_ctx* c, long long* count_out, int reset, void* stream) {
    IVOSW_REQUIRE(c && count_out, "null pointer");
    IVOSW_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    if ((rc = ensure_pinned(c, 64))) return rc;
    IVOSW_CUDA(cudaMemcpyAsync(c->pinned_small, c->sat_count, sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    if (reset) IVOSW_CUDA(cudaMemsetAsync(c->sat_count, 0, sizeof(unsigned long long), s));
    IVOSW_CUDA(cudaStreamSynchronize(s));
    *count_out = (long long)*(unsigned long long*)c->pinned_small;
    if (*count_out > 0) {
        char buf[200];
        snprintf(buf, sizeof buf, "warning: %lld epilogue tiles clamped activations to the fp16 range (+-65504) in the "
                 "split-fp16 encoder; results are outside the parity guarantee", *count_out);
        set_error(buf);
    }
    return IVOSW_OK;
}

int ivosw_dqn_get_optimizer(ivosw_ctx* c, float* m_dev, float* v_dev, long long* step_out, void* stream) {
    IVOSW_REQUIRE(c && m_dev && v_dev && step_out, "null pointer");
    IVOSW_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t nb = sizeof(float) * IVOSW_BRAIN_NUM_PARAMS;
    if (c->adam_m) {
        IVOSW_CUDA(cudaMemcpyAsync(m_dev, c->adam_m, nb, cudaMemcpyDeviceToDevice, s));
        IVOSW_CUDA(cudaMemcpyAsync(v_dev, c->adam_v, nb, cudaMemcpyDeviceToDevice, s));
    } else {
        IVOSW_CUDA(cudaMemsetAsync(m_dev, 0, nb, s));
        IVOSW_CUDA(cudaMemsetAsync(v_dev, 0, nb, s));
    }
    *step_out = c->adam_step;
    return IVOSW_OK;
}

int ivosw_dqn_set_optimizer(ivosw_ctx* c, const float* m_dev, const float* v_dev, long long step, void* stream) {
    IVOSW_REQUIRE(c && m_dev && v_dev && step >= 0, "null pointer / negative step");
    IVOSW_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    const size_t nb = sizeof(float) * IVOSW_BRAIN_NUM_PARAMS;
    if (!c->adam_m) {
        IVOSW_CUDA(cudaMalloc(&c->adam_m, nb));
        IVOSW_CUDA(cudaMalloc(&c->adam_v, nb));
    }
    IVOSW_CUDA(cudaMemcpyAsync(c->adam_m, m_dev, nb, cudaMemcpyDeviceToDevice, s));
    IVOSW_CUDA(cudaMemcpyAsync(c->adam_v, v_dev, nb, cudaMemcpyDeviceToDevice, s));
    c->adam_step = step;
    return IVOSW_OK;
}

// ---------------------------------------------------------------- MANet feature extractor (restatement, parity unpinned)
size_t ivosw_manet_encoder_blob_floats(void) { return manet_encoder_blob_floats(); }

int ivosw_manet_encoder_load(ivosw_ctx* c, const float* blob, size_t n_floats) {
    IVOSW_REQUIRE(c && blob, "null pointer");
    IVOSW_CUDA(cudaSetDevice(c->device));
    return manet_encoder_load(c, blob, n_floats);
}

int ivosw_manet_encoder_forward(ivosw_ctx* c, const float* frames_dev, int B, int H, int W, float* embedding_dev, void* stream) {
    IVOSW_REQUIRE(c && frames_dev && embedding_dev, "null pointer");
    IVOSW_REQUIRE(B >= 1 && H >= 16 && W >= 16, "B, H, W");
    IVOSW_REQUIRE(c->conv_mode != IVOSW_CONV_SIMT_FP32, "the encoder runs on the tensor-core modes only");
    IVOSW_CUDA(cudaSetDevice(c->device));
    const int terms = c->conv_mode == IVOSW_CONV_TC_FP16X3 ? 3 : 1;
    const int chunk = 8;                        // frames per pass (workspace ~ 0.35 GB per 480 x 854 frame)
    const int h4 = (((H + 6 - 7) / 2 + 1) + 2 - 3) / 2 + 1, w4 = (((W + 6 - 7) / 2 + 1) + 2 - 3) / 2 + 1;
    for (int b0 = 0; b0 < B; b0 += chunk) {
        const int nb = std::min(chunk, B - b0);
        int rc = manet_encoder_forward(c, frames_dev + (size_t)b0 * 3 * H * W, nb, H, W, embedding_dev + (size_t)b0 * 100 * h4 * w4,
                                       terms, (cudaStream_t)stream);
        if (rc) return rc;
    }
    return IVOSW_OK;
}

// ---------------------------------------------------------------- AssessNet training step (config C5)
int ivosw_assess_train_begin(ivosw_ctx* c, const float* blob, size_t n_floats) {
    IVOSW_REQUIRE(c && blob, "null pointer");
    IVOSW_REQUIRE(n_floats == assess_blob_floats(), "AssessNet blob length");
    IVOSW_CUDA(cudaSetDevice(c->device));
    for (int i = 0; i < 3; ++i) { c->mean[i] = blob[i]; c->stdv[i] = blob[3 + i]; }
    return train_begin(c, blob, n_floats);
}

int ivosw_assess_train_step(ivosw_ctx* c, const float* frames_dev, long long frame_stride, const float* prob_dev,
                            long long prob_stride, int B, int H, int W, const float* targets_dev, const int* valid_dev,
                            float lr, float momentum, float weight_decay, int apply_update, float* loss_host,
                            float* pred_host, void* stream) {
    IVOSW_REQUIRE(c && frames_dev && prob_dev && targets_dev && valid_dev, "null pointer");
    IVOSW_REQUIRE(B >= 1 && H >= 2 && W >= 2, "B, H, W");
    IVOSW_CUDA(cudaSetDevice(c->device));
    int rc;
    if ((rc = ensure(c->boxes, (size_t)B * 4 * sizeof(float)))) return rc;
    UnitAddr ua{frames_dev, frame_stride, prob_dev, prob_stride, 0, B, 0};
    return train_step(c, ua, B, H, W, targets_dev, valid_dev, lr, momentum, weight_decay, apply_update, loss_host, pred_host,
                      (cudaStream_t)stream);
}

int ivosw_assess_train_apply(ivosw_ctx* c, float lr, float momentum, float weight_decay, void* stream) {
    IVOSW_REQUIRE(c != nullptr, "ctx");
    IVOSW_CUDA(cudaSetDevice(c->device));
    int rc = train_apply(c, lr, momentum, weight_decay, (cudaStream_t)stream);
    if (rc) return rc;
    IVOSW_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return IVOSW_OK;
}

int ivosw_assess_train_export(ivosw_ctx* c, float* blob_host, float* grad_host, void* stream) {
    IVOSW_REQUIRE(c != nullptr, "ctx");
    IVOSW_CUDA(cudaSetDevice(c->device));
    return train_export(c, blob_host, grad_host, (cudaStream_t)stream);
}

int ivosw_assess_train_grads(ivosw_ctx* c, float** grads_dev, size_t* n_floats) {
    IVOSW_REQUIRE(c && grads_dev && n_floats, "null pointer");
    return train_grad_buffer(c, grads_dev, n_floats);
}

// ---------------------------------------------------------------- peer-memory gather (multi-GPU round)
int ivosw_gather_create(ivosw_ctx* c, int world, int rank, int capacity, void* handle_out) {
    IVOSW_REQUIRE(c && handle_out, "null pointer");
    IVOSW_REQUIRE(world >= 1 && world <= 64 && rank >= 0 && rank < world && capacity >= 1, "world (<= 64), rank, capacity");
    IVOSW_CUDA(cudaSetDevice(c->device));
    return gather_create(c, world, rank, capacity, handle_out);
}

int ivosw_gather_open(ivosw_ctx* c, const void* handles) {
    IVOSW_REQUIRE(c && handles, "null pointer");
    IVOSW_CUDA(cudaSetDevice(c->device));
    return gather_open(c, handles);
}

int ivosw_gather_post(ivosw_ctx* c, const double* mq_local_dev, int n_local, int offset, void* stream) {
    IVOSW_REQUIRE(c && (mq_local_dev || n_local == 0), "null pointer");
    IVOSW_CUDA(cudaSetDevice(c->device));
    return gather_post(c, mq_local_dev, n_local, offset, (cudaStream_t)stream);
}

int ivosw_agent_action_gathered(ivosw_ctx* c, const double* ann_host, int T, float* q_host, int* next_frame, double* mq_host,
                                void* stream) {
    IVOSW_REQUIRE(c && ann_host, "null pointer");
    IVOSW_REQUIRE(T >= 1, "T");
    if (!c->brain_loaded) { set_error("Brain weights not loaded"); return IVOSW_ERR_STATE; }
    IVOSW_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    if ((rc = ensure(c->mq, sizeof(double) * 2 * (size_t)T))) return rc;
    if ((rc = ensure(c->brain_state, sizeof(float) * 2 * (size_t)T))) return rc;
    if ((rc = ensure(c->brain_q, sizeof(float) * (size_t)T))) return rc;
    if ((rc = ensure(c->brain_arg, sizeof(int)))) return rc;
    if ((rc = ensure(c->brain_gi, sizeof(float) * (size_t)T * 512))) return rc;
    if ((rc = ensure(c->brain_h, sizeof(float) * (size_t)2 * T * 128))) return rc;
    if ((rc = ensure_pinned(c, sizeof(double) * 2 * T + sizeof(float) * T + 64))) return rc;
    double* pin_ann = (double*)c->pinned_small;
    double* pin_mq = pin_ann + T;
    float* pin_q = (float*)(pin_mq + T);
    int* pin_arg = (int*)(pin_q + T);
    double* mq_dev = (double*)c->mq.p;        // gathered vector, copied out of the gather buffer for the caller
    double* ann_dev = mq_dev + T;
    memcpy(pin_ann, ann_host, sizeof(double) * T);
    auto enqueue = [&](cudaStream_t st) -> int {
        int r;
        IVOSW_CUDA(cudaMemcpyAsync(ann_dev, pin_ann, sizeof(double) * T, cudaMemcpyHostToDevice, st));
        if ((r = gather_wait_pack(c, ann_dev, T, (float*)c->brain_state.p, mq_dev, st))) return r;
        if ((r = timed_brain(c, T, st))) return r;
        IVOSW_CUDA(cudaMemcpyAsync(pin_q, c->brain_q.p, sizeof(float) * T, cudaMemcpyDeviceToHost, st));
        IVOSW_CUDA(cudaMemcpyAsync(pin_arg, c->brain_arg.p, sizeof(int), cudaMemcpyDeviceToHost, st));
        IVOSW_CUDA(cudaMemcpyAsync(pin_mq, mq_dev, sizeof(double) * T, cudaMemcpyDeviceToHost, st));
        return IVOSW_OK;
    };
    ivosw_ctx::GraphKey key{c->gather_state, nullptr, nullptr, T, 0, 0, 0, 0, 0, 0, 6, 0};
    if ((rc = run_graphed(c, key, s, enqueue))) return rc;
    IVOSW_CUDA(cudaStreamSynchronize(s));
    drain_graph_events(c);
    if (q_host) memcpy(q_host, pin_q, sizeof(float) * T);
    if (next_frame) *next_frame = *pin_arg;
    if (mq_host) memcpy(mq_host, pin_mq, sizeof(double) * T);
    return IVOSW_OK;
}

int ivosw_stage_timing(ivosw_ctx* c, int enable) {
    IVOSW_REQUIRE(c != nullptr, "ctx");
    c->timing_on = enable != 0;
    return IVOSW_OK;
}

int ivosw_stage_times(ivosw_ctx* c, float* ms_out, long long* n_conv_launches, int reset) {
    IVOSW_REQUIRE(c && ms_out, "null pointer");
    IVOSW_CUDA(cudaSetDevice(c->device));
    drain_stage_events(c);
    for (int i = 0; i < IVOSW_NUM_STAGES; ++i) ms_out[i] = c->stage_ms[i];
    if (n_conv_launches) *n_conv_launches = c->conv_launches_timed;
    if (reset) {
        for (int i = 0; i < IVOSW_NUM_STAGES; ++i) c->stage_ms[i] = 0.f;
        c->conv_launches_timed = 0;
    }
    return IVOSW_OK;
}

int ivosw_debug_conv(ivosw_ctx* c, int li, int conv_mode, const float* in_dev, const float* residual_dev,
                     float* out_dev, int B, int* dims_out, void* stream) {
    IVOSW_REQUIRE(c != nullptr, "ctx");
    IVOSW_REQUIRE(li >= 0 && li < (int)c->layers.size(), "layer_index");
    IVOSW_REQUIRE(conv_mode >= 0 && conv_mode <= 2, "conv_mode");
    const ConvLayer& L = c->layers[li];
    if (dims_out) {
        dims_out[0] = L.cin; dims_out[1] = L.in_hw; dims_out[2] = L.cout; dims_out[3] = L.out_hw;
        dims_out[4] = L.k; dims_out[5] = L.stride;
    }
    if (!in_dev) return IVOSW_OK;   // query only
    IVOSW_REQUIRE(out_dev && B >= 1, "out_dev, B");
    if (!c->assess_loaded) { set_error("AssessNet weights not loaded"); return IVOSW_ERR_STATE; }
    IVOSW_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (conv_mode == IVOSW_CONV_SIMT_FP32) return launch_conv_simt(c, L, in_dev, residual_dev, out_dev, B, s);
    int rc;
    const size_t n_in = (size_t)B * L.in_hw * L.in_hw * L.cin, n_out = (size_t)B * L.out_hw * L.out_hw * L.cout;
    if ((rc = ensure(c->actX, n_in * 4))) return rc;
    if ((rc = ensure(c->actY, n_out * 4))) return rc;
    if ((rc = ensure(c->actDS, n_out * 4))) return rc;
    const SplitAct xin = split_view(c->actX), yout = split_view(c->actY), res = split_view(c->actDS);
    const int terms = conv_mode == IVOSW_CONV_TC_FP16X3 ? 3 : 1;
    if ((rc = launch_split(c, in_dev, xin, (long long)n_in, s))) return rc;
    if (residual_dev && (rc = launch_split(c, residual_dev, res, (long long)n_out, s))) return rc;
    // measurement only: IVOSW_DEBUG_CONV_REPEAT = n launches the (idempotent) layer n times, so that the
    // difference between two n isolates the convolution kernel from the split / merge helpers
    int reps = 1;
    if (const char* e = getenv("IVOSW_DEBUG_CONV_REPEAT")) reps = atoi(e) > 1 ? atoi(e) : 1;
    for (int r = 0; r < reps; ++r)
        if ((rc = launch_conv_tc(c, L, xin, residual_dev ? &res : nullptr, yout, B, terms, s))) return rc;
    return launch_merge(c, yout, out_dev, (long long)n_out, terms == 3, s);
}

// -------------------------------------------------------------------------------------- MANet tail
int ivosw_manet_tail(ivosw_ctx* c, const float* logits_dev, int T, int C, int h, int w, int H, int W,
                     float* masks_dev, float* all_p_dev, void* stream) {
    IVOSW_REQUIRE(c && logits_dev, "null pointer");
    IVOSW_REQUIRE(T >= 1 && C >= 1 && C <= 16 && h >= 1 && w >= 1 && H >= 1 && W >= 1, "T, C (<= 16), h, w, H, W");
    IVOSW_REQUIRE(T <= 65535 && H <= 65535, "T, H <= 65535");
    IVOSW_CUDA(cudaSetDevice(c->device));
    return launch_manet_tail(c, logits_dev, T, C, h, w, H, W, masks_dev, all_p_dev, (cudaStream_t)stream);
}

int ivosw_rough_roi(ivosw_ctx* c, const float* labels_dev, float* out_dev, int B, int h, int w, int dist,
                    void* stream) {
    IVOSW_REQUIRE(c && labels_dev && out_dev, "null pointer");
    IVOSW_REQUIRE(B >= 1 && h >= 1 && w >= 1 && dist >= 0, "B, h, w, dist");
    IVOSW_CUDA(cudaSetDevice(c->device));
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    if ((rc = ensure(c->brain_arg, sizeof(int)))) return rc;
    if ((rc = ensure_pinned(c, 64))) return rc;
    if ((rc = launch_rough_roi(c, labels_dev, out_dev, B, h, w, dist, (int*)c->brain_arg.p, s))) return rc;
    IVOSW_CUDA(cudaMemcpyAsync(c->pinned_small, c->brain_arg.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    IVOSW_CUDA(cudaStreamSynchronize(s));
    if (*(int*)c->pinned_small) {
        set_error("rough_ROI: an image has no pixel different from -1 (the reference raises on the empty min/max)");
        return IVOSW_ERR_INVALID;
    }
    return IVOSW_OK;
}

// -------------------------------------------------------------------------------------- ATNet glue
int ivosw_atnet_reflect_pad(ivosw_ctx* c, const float* in_dev, float* out_dev, int planes, int h, int w, int left, int right,
                            int top, int bottom, void* stream) {
    IVOSW_REQUIRE(c && in_dev && out_dev, "null pointer");
    IVOSW_REQUIRE(planes >= 1 && planes <= 65535 && h >= 1 && w >= 1, "planes, h, w");
    // torch.nn.ReflectionPad2d: "Padding size should be less than the corresponding input dimension"
    IVOSW_REQUIRE(left >= 0 && right >= 0 && top >= 0 && bottom >= 0 && left < w && right < w && top < h && bottom < h,
                  "padding must be non-negative and smaller than the input dimension");
    IVOSW_REQUIRE(h + top + bottom <= 65535, "padded height <= 65535");
    IVOSW_CUDA(cudaSetDevice(c->device));
    return launch_reflect_pad(c, in_dev, out_dev, planes, h, w, left, right, top, bottom, (cudaStream_t)stream);
}

int ivosw_atnet_sigmoid_blend(ivosw_ctx* c, const float* logit_dev, const float* prev_dev, float* prob_dev, float* blended_dev,
                              long long n, float alpha, float one_minus_alpha, void* stream) {
    IVOSW_REQUIRE(c && logit_dev && prob_dev && blended_dev, "null pointer");
    IVOSW_REQUIRE(n >= 1 && n < (1ll << 40), "n");
    IVOSW_CUDA(cudaSetDevice(c->device));
    return launch_sigmoid_blend(c, logit_dev, prev_dev, prob_dev, blended_dev, n, alpha, one_minus_alpha, (cudaStream_t)stream);
}

int ivosw_atnet_assemble(ivosw_ctx* c, const float* prob_map_dev, float* all_p_dev, int T, int O, int PH, int PW, int y0, int x0,
                         int H, int W, void* stream) {
    IVOSW_REQUIRE(c && prob_map_dev && all_p_dev, "null pointer");
    IVOSW_REQUIRE(T >= 1 && O >= 1 && H >= 1 && W >= 1 && y0 >= 0 && x0 >= 0 && y0 + H <= PH && x0 + W <= PW, "geometry");
    IVOSW_REQUIRE((long long)T * (O + 1) <= 65535 && H <= 65535, "T*(O+1), H <= 65535");
    IVOSW_CUDA(cudaSetDevice(c->device));
    return launch_atnet_assemble(c, prob_map_dev, all_p_dev, T, O, PH, PW, y0, x0, H, W, (cudaStream_t)stream);
}

}  // extern "C"

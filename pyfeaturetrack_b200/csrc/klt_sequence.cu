// klt_sequence: B lock-stepped frame sequences in sequentialMode with per-frame feature replacement -- BASELINE config D.
//
// Reference flow per frame (one sequence): KLTTrackFeatures(tc, prev, cur, fl) with tc.sequentialMode (pyramid reuse,
// trackFeatures.py:152-161,401-404), then KLTReplaceLostFeatures(tc, cur, fl) = _KLTSelectGoodFeatures(REPLACING_SOME) on the
// gradients of tc.pyramid_last (selectGoodFeatures.py:176-179, 45-135).  Here one step does that for B independent sequences at
// once: ONE pyramid build (the new frames), ONE tracking launch, ONE selection chain, all on device-resident feature lists,
// with no host synchronisation anywhere; from the third step on the whole chain is replayed from a CUDA graph (one graph per
// pyramid parity).  Host frames are uploaded on a second stream into two alternating staging buffers, so the upload of step k+1
// overlaps the kernels of step k when the caller does not wait in between.
#include <cstdlib>
#include <cstring>

#include "klt_common.cuh"
#include "klt_select.cuh"

struct klt_sequence {
    klt_params params;
    klt_taps taps;
    int w, h, B, n, precision, select_mode;
    klt_pyr *pyr[2];
    int cur;                         // pyramid that holds the latest frame
    int started;
    long steps;
    SelDev S_all, S_rep;
    float *sat;
    bool fast_select_ok;
    char *block;                     // selection workspace + feature lists + counters (one allocation)
    double *fx, *fy;
    int *fval, *fval_tracked;
    unsigned long long *iters;       // [0] Newton iterations since the last klt_sequence_sync
    int *aflag;                      // sticky "a window left the image" flag (the reference's AssertionError case)
    uint8_t *stage[2];
    size_t stage_bytes, cap_pitch, cap_stride;
    cudaEvent_t stage_free[2], stage_ready[2];
    cudaEvent_t fork_ev, join_ev;    // the eigenvalue pass of a step runs beside the tracking kernel on the context's second stream
    int overlap;
    int use_graph;
    cudaGraphExec_t graph[2][2];     // [pyramid parity][replace]
    bool graph_ok[2][2];
    int warm[2][2];
    int64_t graph_launches[2][2];
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// eigenvalue maps of the level-0 planes of `p` (all images): the part of a selection that does not depend on the feature lists
static int eigen_on(klt_ctx *ctx, klt_sequence *q, klt_pyr *p, const SelDev *S) {
    int rc, done = 0;
    if (q->select_mode == KLT_SELECT_FAST && q->fast_select_ok) {
        done = klt_sel_launch_eigen_fast(ctx, S, q->B, p->level(0, 0, 0), p->plane_floats, p->lv[0].pitch, &q->taps.grad_gauss,
                                         &q->taps.grad_deriv);
        if (done < 0) return done;
    }
    if (!done) {
        if ((rc = klt_ensure_gradients_level0(ctx, p))) return rc;
        if ((rc = klt_sel_launch_eigen_strict(ctx, S, q->B, p->level(1, 0, 0), p->level(2, 0, 0), p->plane_floats, p->lv[0].pitch,
                                              q->sat))) return rc;
    }
    return KLT_OK;
}

// selection on the level-0 planes of `p` (all images) into the sequence's feature lists
static int select_on(klt_ctx *ctx, klt_sequence *q, klt_pyr *p, const SelDev *S) {
    int rc;
    if ((rc = klt_sel_launch_begin(ctx, S, q->B))) return rc;
    if ((rc = eigen_on(ctx, q, p, S))) return rc;
    return klt_sel_launch_pick(ctx, S, q->B);
}

// everything of one step that runs on the compute stream after the frames are in stage[par]
static int enqueue_step(klt_ctx *ctx, klt_sequence *q, int par, int replace) {
    const int prev = q->cur, cur = par;
    bool windowed;
    const int arith = klt_begin_build(q->pyr[cur], &q->taps, q->precision, &windowed);
    int rc;
    // The eigenvalue maps need the new pyramid only -- the fused fast pass just its level-0 image -- while the rest of the
    // replacement (pre-marking, histogram, walk) needs the tracked lists: the map pass runs on the second stream beside the
    // decimations and the tracking kernel (all issue-bound or small; together they fill the SMs better than one after the
    // other).  Profiling runs keep one stream so that per-kernel times stay meaningful.
    const bool fork = replace && q->overlap && !ctx->profiling && ctx->aux_stream;
    const bool fork_early = fork && q->select_mode == KLT_SELECT_FAST && q->fast_select_ok;    // forks right after the level-0 kernel
    ctx->level0_event = fork_early ? q->fork_ev : nullptr;
    rc = klt_build_u8_device(ctx, q->pyr[cur], q->stage[par], q->cap_pitch, q->cap_stride, &q->taps, arith, 0, q->B, windowed);
    ctx->level0_event = nullptr;
    if (rc) return rc;
    if (fork) {
        cudaStream_t main = ctx->stream;
        if (!fork_early) KLT_CUDA(ctx, cudaEventRecord(q->fork_ev, main));
        KLT_CUDA(ctx, cudaStreamWaitEvent(ctx->aux_stream, q->fork_ev, 0));
        ctx->stream = ctx->aux_stream;
        rc = eigen_on(ctx, q, q->pyr[cur], &q->S_rep);
        ctx->stream = main;
        if (rc) return rc;
        KLT_CUDA(ctx, cudaEventRecord(q->join_ev, ctx->aux_stream));
    }
    if ((rc = klt_launch_track(ctx, &q->params, q->pyr[prev], q->pyr[cur], q->n, q->fx, q->fy, q->fval, q->iters, q->aflag))) return rc;
    KLT_CUDA(ctx, cudaMemcpyAsync(q->fval_tracked, q->fval, (size_t)q->B * q->n * sizeof(int), cudaMemcpyDeviceToDevice, ctx->stream));
    if (fork) {
        if ((rc = klt_sel_launch_begin(ctx, &q->S_rep, q->B))) return rc;
        KLT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, q->join_ev, 0));
        return klt_sel_launch_pick(ctx, &q->S_rep, q->B);
    }
    if (replace && (rc = select_on(ctx, q, q->pyr[cur], &q->S_rep))) return rc;
    return KLT_OK;
}

static int stage_frames(klt_ctx *ctx, klt_sequence *q, int par, const uint8_t *frames, size_t pitch, size_t frame_stride) {
    if (pitch < (size_t)q->w) return klt_fail(ctx, KLT_ERR_INVALID, "pitch %zu smaller than width %d", pitch, q->w);
    const size_t bytes = (size_t)(q->B - 1) * frame_stride + (size_t)(q->h - 1) * pitch + q->w;
    if (bytes > q->stage_bytes) {
        KLT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        KLT_CUDA(ctx, cudaStreamSynchronize(ctx->copy_stream));
        for (int i = 0; i < 2; i++) {
            if (q->stage[i]) KLT_CUDA(ctx, cudaFree(q->stage[i]));
            q->stage[i] = nullptr;
            KLT_CUDA(ctx, cudaMalloc(&q->stage[i], bytes + 256));
        }
        q->stage_bytes = bytes;
    }
    if (pitch != q->cap_pitch || frame_stride != q->cap_stride) {      // captured graphs carry the frame layout
        for (int a = 0; a < 2; a++)
            for (int r = 0; r < 2; r++) {
                if (q->graph_ok[a][r]) { cudaGraphExecDestroy(q->graph[a][r]); q->graph_ok[a][r] = false; }
                q->warm[a][r] = 0;
            }
        q->cap_pitch = pitch; q->cap_stride = frame_stride;
    }
    if (klt_is_device_ptr(frames)) {
        KLT_CUDA(ctx, cudaMemcpyAsync(q->stage[par], frames, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
    } else {
        KLT_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, q->stage_free[par], 0));     // the build of two steps ago has read it
        KLT_CUDA(ctx, cudaMemcpyAsync(q->stage[par], frames, bytes, cudaMemcpyHostToDevice, ctx->copy_stream));
        KLT_CUDA(ctx, cudaEventRecord(q->stage_ready[par], ctx->copy_stream));
        KLT_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, q->stage_ready[par], 0));
    }
    return KLT_OK;
}

extern "C" {

int klt_sequence_create(klt_ctx *ctx, const klt_params *params, const klt_taps *taps, int w, int h, int n_sequences,
                        int n_features, int precision, int select_mode, klt_sequence **out) {
    if (!ctx || !params || !taps || !out || w <= 0 || h <= 0 || n_sequences <= 0 || n_features <= 0)
        return klt_fail(ctx, KLT_ERR_INVALID, "bad argument");
    if (select_mode != KLT_SELECT_STRICT && select_mode != KLT_SELECT_FAST) return klt_fail(ctx, KLT_ERR_INVALID, "select_mode must be KLT_SELECT_STRICT or KLT_SELECT_FAST");
    if (precision < 0 || precision > 2) return klt_fail(ctx, KLT_ERR_INVALID, "bad precision");
    if (params->affine_consistency_check >= 0 || params->lighting_insensitive)
        return klt_fail(ctx, KLT_ERR_UNSUPPORTED, "klt_sequence covers translational tracking; use klt_track_features_affine per frame for the affine check");
    KLT_CUDA(ctx, cudaSetDevice(ctx->device));
    *out = nullptr;
    klt_sequence *q = new klt_sequence();
    memset(q, 0, sizeof *q);
    q->params = *params; q->taps = *taps;
    q->w = w; q->h = h; q->B = n_sequences; q->n = n_features; q->precision = precision; q->select_mode = select_mode;
    q->use_graph = getenv("KLT_B200_NO_GRAPH") ? 0 : 1;
    q->overlap = (getenv("KLT_B200_SEQ_OVERLAP") && atoi(getenv("KLT_B200_SEQ_OVERLAP")) == 0) ? 0 : 1;
    int rc;
    for (int i = 0; i < 2; i++)
        if ((rc = klt_pyr_create(ctx, w, h, params->n_levels, params->subsampling, n_sequences, &q->pyr[i]))) { klt_sequence_destroy(ctx, q); return rc; }
    if ((rc = klt_sel_geometry(ctx, params, w, h, n_features, 1, &q->S_rep))) { klt_sequence_destroy(ctx, q); return rc; }
    // the fused fast eigen pass covers the default gradient kernel (7 taps) and square windows up to 15
    q->fast_select_ok = select_mode == KLT_SELECT_FAST && taps->grad_gauss.n == 7 && taps->grad_deriv.n == 7 &&
                        q->S_rep.hw == q->S_rep.hh && q->S_rep.hw >= 1 && q->S_rep.hw <= 7;
    const bool need_sat = !q->fast_select_ok;
    const size_t total = (size_t)q->B * q->n;
    const size_t sel_b = klt_sel_workspace_bytes(&q->S_rep, q->B, need_sat, true);
    const size_t extra = align_up(total * sizeof(int), 256) + 512;
    cudaError_t e = cudaMalloc(&q->block, sel_b + extra);
    if (e != cudaSuccess) { klt_sequence_destroy(ctx, q); return klt_fail(ctx, KLT_ERR_NOMEM, "cudaMalloc(%zu) for the sequence workspace failed: %s", sel_b + extra, cudaGetErrorString(e)); }
    klt_sel_carve(&q->S_rep, q->B, need_sat, true, q->block, &q->sat);
    q->fx = q->S_rep.fx; q->fy = q->S_rep.fy; q->fval = q->S_rep.fval;
    q->fval_tracked = (int *)(q->block + sel_b);
    q->iters = (unsigned long long *)(q->block + sel_b + align_up(total * sizeof(int), 256));
    q->aflag = (int *)(q->iters + 4);
    cudaMemsetAsync(q->block + sel_b, 0, extra, ctx->stream);
    q->S_all = q->S_rep;                    // same buffers, SELECTING_ALL semantics
    q->S_all.replace = 0; q->S_all.premap = nullptr; q->S_all.target_mul = 32u; q->S_all.target_add = 8192u;
    if ((rc = klt_sel_prepare_kernels(ctx, &q->S_rep))) { klt_sequence_destroy(ctx, q); return rc; }
    for (int i = 0; i < 2; i++) {
        cudaEventCreateWithFlags(&q->stage_free[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&q->stage_ready[i], cudaEventDisableTiming);
    }
    cudaEventCreateWithFlags(&q->fork_ev, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&q->join_ev, cudaEventDisableTiming);
    *out = q;
    return KLT_OK;
}

int klt_sequence_destroy(klt_ctx *ctx, klt_sequence *q) {
    if (!q) return KLT_OK;
    if (ctx) { cudaStreamSynchronize(ctx->stream); cudaStreamSynchronize(ctx->copy_stream); }
    for (int a = 0; a < 2; a++)
        for (int r = 0; r < 2; r++)
            if (q->graph_ok[a][r]) cudaGraphExecDestroy(q->graph[a][r]);
    for (int i = 0; i < 2; i++) {
        if (q->pyr[i]) klt_pyr_destroy(ctx, q->pyr[i]);
        if (q->stage[i]) cudaFree(q->stage[i]);
        if (q->stage_free[i]) cudaEventDestroy(q->stage_free[i]);
        if (q->stage_ready[i]) cudaEventDestroy(q->stage_ready[i]);
    }
    if (q->fork_ev) cudaEventDestroy(q->fork_ev);
    if (q->join_ev) cudaEventDestroy(q->join_ev);
    if (q->block) cudaFree(q->block);
    delete q;
    return KLT_OK;
}

int klt_sequence_start_u8(klt_ctx *ctx, klt_sequence *q, const uint8_t *frames, size_t pitch, size_t frame_stride, int select) {
    if (!ctx || !q || !frames) return klt_fail(ctx, KLT_ERR_INVALID, "bad argument");
    KLT_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    const int par = 0;
    if ((rc = stage_frames(ctx, q, par, frames, pitch, frame_stride))) return rc;
    bool windowed;
    const int arith = klt_begin_build(q->pyr[par], &q->taps, q->precision, &windowed);
    if ((rc = klt_build_u8_device(ctx, q->pyr[par], q->stage[par], pitch, frame_stride, &q->taps, arith, 0, q->B, windowed))) return rc;
    KLT_CUDA(ctx, cudaEventRecord(q->stage_free[par], ctx->stream));
    if (select && (rc = select_on(ctx, q, q->pyr[par], &q->S_all))) return rc;
    q->cur = par; q->started = 1;
    return KLT_OK;
}

int klt_sequence_set_features(klt_ctx *ctx, klt_sequence *q, const double *x, const double *y, const int32_t *val) {
    if (!ctx || !q || !x || !y || !val) return klt_fail(ctx, KLT_ERR_INVALID, "bad argument");
    const size_t total = (size_t)q->B * q->n;
    KLT_CUDA(ctx, cudaMemcpyAsync(q->fx, x, total * sizeof(double), cudaMemcpyDefault, ctx->stream));
    KLT_CUDA(ctx, cudaMemcpyAsync(q->fy, y, total * sizeof(double), cudaMemcpyDefault, ctx->stream));
    KLT_CUDA(ctx, cudaMemcpyAsync(q->fval, val, total * sizeof(int32_t), cudaMemcpyDefault, ctx->stream));
    return KLT_OK;
}

static int copy_out(klt_ctx *ctx, klt_sequence *q, double *x, double *y, int32_t *val, int32_t *val_tracked) {
    const size_t total = (size_t)q->B * q->n;
    if (x) KLT_CUDA(ctx, cudaMemcpyAsync(x, q->fx, total * sizeof(double), cudaMemcpyDefault, ctx->stream));
    if (y) KLT_CUDA(ctx, cudaMemcpyAsync(y, q->fy, total * sizeof(double), cudaMemcpyDefault, ctx->stream));
    if (val) KLT_CUDA(ctx, cudaMemcpyAsync(val, q->fval, total * sizeof(int32_t), cudaMemcpyDefault, ctx->stream));
    if (val_tracked) KLT_CUDA(ctx, cudaMemcpyAsync(val_tracked, q->fval_tracked, total * sizeof(int32_t), cudaMemcpyDefault, ctx->stream));
    return KLT_OK;
}

int klt_sequence_step_u8(klt_ctx *ctx, klt_sequence *q, const uint8_t *frames, size_t pitch, size_t frame_stride, int replace,
                         double *x, double *y, int32_t *val, int32_t *val_tracked) {
    if (!ctx || !q || !frames) return klt_fail(ctx, KLT_ERR_INVALID, "bad argument");
    if (!q->started) return klt_fail(ctx, KLT_ERR_INVALID, "klt_sequence_start_u8 has not been called");
    KLT_CUDA(ctx, cudaSetDevice(ctx->device));
    int rc;
    const int par = q->cur ^ 1, rep = replace ? 1 : 0;
    if ((rc = stage_frames(ctx, q, par, frames, pitch, frame_stride))) return rc;
    bool launched = false;
    if (q->use_graph && !ctx->profiling) {
        if (!q->graph_ok[par][rep] && q->warm[par][rep]) {
            // capture this step once; the chain of launches is the same for every later step of this parity
            const int64_t l0 = ctx->launches;
            cudaGraph_t g = nullptr;
            if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                rc = enqueue_step(ctx, q, par, rep);
                const cudaError_t ee = cudaStreamEndCapture(ctx->stream, &g);
                if (rc == KLT_OK && ee == cudaSuccess && g && cudaGraphInstantiate(&q->graph[par][rep], g, 0) == cudaSuccess) {
                    q->graph_ok[par][rep] = true;
                    q->graph_launches[par][rep] = ctx->launches - l0;
                } else {
                    cudaGetLastError();
                    q->use_graph = 0;                      // this configuration cannot be captured: plain launches from now on
                }
                if (g) cudaGraphDestroy(g);
                ctx->launches = l0;
            } else {
                cudaGetLastError();
                q->use_graph = 0;
            }
        }
        if (q->graph_ok[par][rep]) {
            // host-side bookkeeping the captured calls would have done
            bool windowed;
            klt_begin_build(q->pyr[par], &q->taps, q->precision, &windowed);
            if (rep && !(q->select_mode == KLT_SELECT_FAST && q->fast_select_ok) && q->pyr[par]->hx) q->pyr[par]->hx->grad0_valid = true;
            KLT_CUDA(ctx, cudaGraphLaunch(q->graph[par][rep], ctx->stream));
            ctx->launches += q->graph_launches[par][rep];
            launched = true;
        }
    }
    if (!launched) {
        if ((rc = enqueue_step(ctx, q, par, rep))) return rc;
        q->warm[par][rep] = 1;
    }
    KLT_CUDA(ctx, cudaEventRecord(q->stage_free[par], ctx->stream));
    q->cur = par;
    q->steps++;
    return copy_out(ctx, q, x, y, val, val_tracked);
}

int klt_sequence_get_features(klt_ctx *ctx, klt_sequence *q, double *x, double *y, int32_t *val, int32_t *val_tracked) {
    if (!ctx || !q) return klt_fail(ctx, KLT_ERR_INVALID, "bad argument");
    int rc = copy_out(ctx, q, x, y, val, val_tracked);
    if (rc) return rc;
    KLT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KLT_OK;
}

int klt_sequence_sync(klt_ctx *ctx, klt_sequence *q, int64_t *n_iterations) {
    if (!ctx || !q) return klt_fail(ctx, KLT_ERR_INVALID, "bad argument");
    unsigned long long res[5] = {0, 0, 0, 0, 0};
    KLT_CUDA(ctx, cudaMemcpyAsync(res, q->iters, sizeof res, cudaMemcpyDeviceToHost, ctx->stream));
    KLT_CUDA(ctx, cudaMemsetAsync(q->iters, 0, sizeof res, ctx->stream));
    KLT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (n_iterations) *n_iterations = (int64_t)res[0];
    if ((int)(res[4] & 0xffffffffull))
        return klt_fail(ctx, KLT_ERR_ASSERT, "a feature window left the image at a pyramid level: the reference raises AssertionError (trackFeaturesUtils.pyx:35)");
    return KLT_OK;
}

int klt_sequence_select_stats(klt_ctx *ctx, klt_sequence *q, int64_t *out) {
    if (!ctx || !q || !out) return klt_fail(ctx, KLT_ERR_INVALID, "bad argument");
    KLT_CUDA(ctx, cudaMemcpyAsync(out, q->S_rep.status, (size_t)q->B * SEL_STATUS_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    KLT_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return KLT_OK;
}

int klt_sequence_pyramid(klt_sequence *q, klt_pyr **out) {
    if (!q || !out) return KLT_ERR_INVALID;
    *out = q->pyr[q->cur];
    return KLT_OK;
}

int klt_sequence_uses_graph(const klt_sequence *q) {
    if (!q) return 0;
    return (q->graph_ok[0][0] || q->graph_ok[0][1] || q->graph_ok[1][0] || q->graph_ok[1][1]) ? 1 : 0;
}

}  // extern "C"

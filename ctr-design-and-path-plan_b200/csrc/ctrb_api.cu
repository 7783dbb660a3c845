// ctrb_api.cu -- the C ABI of libctrb200.so (include/ctrb200.h): handles, staging of host
// buffers, kernel launch sequencing.  There is deliberately no CPU fallback anywhere: without
// a CUDA device every compute entry point fails with CTRB_ERR_NO_DEVICE / CTRB_ERR_CUDA.
#include <cfloat>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <algorithm>

#include "ctrb_internal.cuh"

namespace ctrb {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int ctx_scratch(ctrb_ctx* ctx, int slot, size_t bytes, void** out) {
    if (bytes == 0) bytes = 16;
    if (ctx->scratch_bytes[slot] < bytes) {
        if (ctx->scratch[slot]) {
            // stream-ordered: earlier kernels on ctx->stream may still read the old buffer
            CTRB_CUDA(cudaStreamSynchronize(ctx->stream));
            CTRB_CUDA(cudaStreamSynchronize(ctx->stream2));
            CTRB_CUDA(cudaFree(ctx->scratch[slot]));
            ctx->scratch[slot] = nullptr;
            ctx->scratch_bytes[slot] = 0;
        }
        size_t cap = bytes + bytes / 4;
        cudaError_t e = cudaMalloc(&ctx->scratch[slot], cap);
        if (e != cudaSuccess) {
            set_error("cudaMalloc(%zu) failed: %s", cap, cudaGetErrorString(e));
            return CTRB_ERR_ALLOC;
        }
        ctx->scratch_bytes[slot] = cap;
    }
    *out = ctx->scratch[slot];
    return CTRB_OK;
}

// launchers implemented in ctrb_model.cu / ctrb_collide.cu
int launch_table(ctrb_ctx* ctx, const ctrb_model* m);
int launch_gcurve(ctrb_ctx* ctx, const ctrb_model* m, long long B, const int32_t* idx, const double* theta, int full,
                  int precision, const long long* offsets, void* out);
int launch_calc_indices(ctrb_ctx* ctx, const ctrb_model* m, long long B, const double* prev_ins, const double* next_ins,
                        int32_t* prev_idx, int32_t* new_idx);
int launch_offsets(ctrb_ctx* ctx, const ctrb_model* m, long long B, const int32_t* idx, int full, long long* offsets);
int launch_eta_ftl(ctrb_ctx* ctx, const ctrb_model* m, long long B, const int32_t* prev_idx, const int32_t* new_idx,
                   const double* velocity, const double* delta_theta, const double* prev_g, const long long* prev_off, double* eta_out,
                   double* ftl_out, double* ftl_mag, int* err_flag);
int launch_eval_fused(ctrb_ctx* ctx, const ctrb_model* m, const ctrb_env* env, const ctrb_cost_desc* cd, long long B,
                      const int32_t* idx, const double* theta, const double* rad, double* obstacle_min,
                      double* goal_dist, double* cost, uint8_t* collide);
int launch_eval_curves(ctrb_ctx* ctx, const ctrb_env* env, long long B, int tube_num, const double* g,
                       const long long* offsets, int max_nodes, const double* rad, double* obstacle_min,
                       double* goal_dist);
int launch_argmin(ctrb_ctx* ctx, long long B, const double* cost, unsigned long long* key_out);
int launch_argmin_finish(ctrb_ctx* ctx, const unsigned long long* key, long long index_base, long long* out);
int launch_eval_curves_cost(ctrb_ctx* ctx, const ctrb_env* env, const ctrb_cost_desc* cd, long long B, int tube_num,
                            const double* g, const long long* offsets, int max_nodes, const double* rad,
                            double* obstacle_min, double* goal_dist, double* cost, uint8_t* collide, int pos_only);
int launch_fma_peak(ctrb_ctx* ctx, int precision, double* tflops);
int static_model_init(const ctrb_model_desc* d, const ModelConst& mc, StModel* out);
int launch_static_tables(ctrb_ctx* ctx, ctrb_model* m);
int launch_static_solve(ctrb_ctx* ctx, const StModel& sm, long long B, const int32_t* idx, const double* theta,
                        const double* x0, int x0_per_config, double xtol, const long long* offsets, double* g_out,
                        double* sol_out, int32_t* nfev_out, int32_t* info_out, int pos_only);
int launch_static_curves(ctrb_ctx* ctx, const StModel& sm, long long B, const int32_t* idx, const double* theta,
                         const double* sol, const long long* offsets, double* g_out, int pos_only);
int launch_static_residual(ctrb_ctx* ctx, const StModel& sm, long long B, const int32_t* idx, const double* theta,
                           const double* q, double* F);

}  // namespace ctrb

using namespace ctrb;

// ------------------------------------------------------------------ staging helpers for CTRB_MEM_HOST
namespace {
struct Stage {
    ctrb_ctx* ctx;
    int slot;
    int rc;
    explicit Stage(ctrb_ctx* c) : ctx(c), slot(1), rc(CTRB_OK) {}
    // device buffer holding a copy of host data
    template <typename T>
    const T* in(const T* host, size_t count, int mem) {
        if (mem == CTRB_MEM_DEVICE || rc) return host;
        void* d = nullptr;
        rc = ctx_scratch(ctx, slot++, count * sizeof(T), &d);
        if (rc) return nullptr;
        if (cudaMemcpyAsync(d, host, count * sizeof(T), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
            set_error("H2D copy failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = CTRB_ERR_CUDA;
        }
        return (const T*)d;
    }
    template <typename T>
    T* out(T* host, size_t count, int mem) {
        if (mem == CTRB_MEM_DEVICE || rc || host == nullptr) return host;
        void* d = nullptr;
        rc = ctx_scratch(ctx, slot++, count * sizeof(T), &d);
        return (T*)d;
    }
    template <typename T>
    void back(T* host, const T* dev, size_t count, int mem) {
        if (mem == CTRB_MEM_DEVICE || rc || host == nullptr) return;
        if (cudaMemcpyAsync(host, dev, count * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) {
            set_error("D2H copy failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = CTRB_ERR_CUDA;
        }
    }
    int finish(int mem) {
        if (rc) return rc;
        if (mem == CTRB_MEM_HOST) {
            CTRB_CUDA(cudaStreamSynchronize(ctx->stream));
        }
        return CTRB_OK;
    }
};

// idx[b][t] in [0, npts[t]) (ValueError in the reference's mirror); `first` = index of idx[0] in the caller's batch
int check_idx_host(const ctrb_model* m, int64_t B, const int32_t* idx, int64_t first = 0) {
    const int T = m->mc.T;
    int bad = 0;
    for (int64_t b = 0; b < B; ++b)
        for (int t = 0; t < T; ++t) bad |= (idx[b * T + t] < 0) | (idx[b * T + t] >= m->mc.npts[t]);
    if (!bad) return CTRB_OK;
    for (int64_t b = 0; b < B; ++b)
        for (int t = 0; t < T; ++t)
            CTRB_REQUIRE(idx[b * T + t] >= 0 && idx[b * T + t] < m->mc.npts[t], "idx[%lld]=%d outside [0,%d)",
                         (long long)((first + b) * T + t), idx[b * T + t], m->mc.npts[t]);
    return CTRB_OK;
}

// Host-memory calls on large batches run as a pipeline over chunks of the batch: while chunk c computes on ctx->stream,
// chunk c+1's inputs are validated on the CPU and copied in on ctx->stream2, and chunk c-1's results are copied out
// behind it.  Two sets of device buffers (scratch slots 1.. and 17..) alternate.  Order on stream2:
//     in(0) in(1) [k(0)] out(0) in(2) [k(1)] out(1) ...      ([k(c)] = wait for chunk c's kernels)
// so in(c+2) never overwrites the inputs that the kernels of chunk c still read, and the kernels of chunk c+2 start
// after in(c+2), i.e. after out(c) has read the outputs they overwrite.
// Chunks stay large (an eighth of the batch, at least 2^17 configurations): a launch over fewer configurations loses
// more to its tail than the overlap gains.
constexpr int64_t kPipeMinBatch = 1 << 19;
inline int64_t pipe_chunk(int64_t B) {
    if (B < kPipeMinBatch) return B;
    int64_t ch = (B + 7) / 8;
    ch = std::min<int64_t>(std::max<int64_t>(ch, 1 << 17), 1 << 22);
    return (ch + 31) / 32 * 32;
}
template <class In, class Launch, class Out>
int host_pipeline(ctrb_ctx* ctx, int64_t B, int64_t CH, In copy_in, Launch launch, Out copy_out) {
    int rc = CTRB_OK;
    int64_t prev0 = 0, prevn = 0;
    int c = 0;
    for (int64_t b0 = 0; b0 < B && rc == CTRB_OK; b0 += CH, ++c) {
        const int64_t nb = std::min(CH, B - b0);
        const int p = c & 1;
        rc = copy_in(b0, nb, p, ctx->stream2);
        if (rc) break;
        if (cudaEventRecord(ctx->ev_chunk[p], ctx->stream2) != cudaSuccess ||
            cudaStreamWaitEvent(ctx->stream, ctx->ev_chunk[p], 0) != cudaSuccess) { rc = CTRB_ERR_CUDA; break; }
        rc = launch(b0, nb, p);
        if (rc) break;
        if (cudaEventRecord(ctx->ev_done[p], ctx->stream) != cudaSuccess) { rc = CTRB_ERR_CUDA; break; }
        if (c > 0) {
            if (cudaStreamWaitEvent(ctx->stream2, ctx->ev_done[p ^ 1], 0) != cudaSuccess) { rc = CTRB_ERR_CUDA; break; }
            rc = copy_out(prev0, prevn, p ^ 1, ctx->stream2);
        }
        prev0 = b0;
        prevn = nb;
    }
    if (rc == CTRB_OK && c > 0) {
        const int p = (c - 1) & 1;
        if (cudaStreamWaitEvent(ctx->stream2, ctx->ev_done[p], 0) != cudaSuccess) rc = CTRB_ERR_CUDA;
        else rc = copy_out(prev0, prevn, p, ctx->stream2);
    }
    const cudaError_t e1 = cudaStreamSynchronize(ctx->stream), e2 = cudaStreamSynchronize(ctx->stream2);
    if (rc == CTRB_OK && (e1 != cudaSuccess || e2 != cudaSuccess)) {
        set_error("host pipeline: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
        rc = CTRB_ERR_CUDA;
    }
    return rc;
}
#define CTRB_COPY(dst, src, bytes, kind, stream)                                                           \
    do {                                                                                                   \
        if ((bytes) && cudaMemcpyAsync(dst, src, bytes, kind, stream) != cudaSuccess) {                    \
            set_error("copy failed: %s", cudaGetErrorString(cudaGetLastError()));                          \
            return (int)CTRB_ERR_CUDA;                                                                     \
        }                                                                                                  \
    } while (0)
}  // namespace

extern "C" {

const char* ctrb_last_error(void) { return g_err; }
const char* ctrb_version(void) { return "ctrb200 0.1 (sm_100a)"; }

int ctrb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

// ------------------------------------------------------------------ ctx
int ctrb_ctx_create(int device, ctrb_ctx** out) {
    CTRB_REQUIRE(out != nullptr, "ctrb_ctx_create: out is NULL");
    *out = nullptr;
    int n = ctrb_device_count();
    if (n <= 0) {
        set_error("no CUDA device: libctrb200 has no CPU fallback");
        return CTRB_ERR_NO_DEVICE;
    }
    CTRB_REQUIRE(device >= 0 && device < n, "device %d out of range (0..%d)", device, n - 1);
    CTRB_CUDA(cudaSetDevice(device));
    ctrb_ctx* c = new ctrb_ctx();
    std::memset(c, 0, sizeof *c);
    c->device = device;
    CTRB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = c->stream;
    CTRB_CUDA(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CTRB_CUDA(cudaEventCreate(&c->ev[i]));
        CTRB_CUDA(cudaEventCreateWithFlags(&c->ev_chunk[i], cudaEventDisableTiming));
        CTRB_CUDA(cudaEventCreateWithFlags(&c->ev_done[i], cudaEventDisableTiming));
    }
    for (int k = 0; k < 8; ++k)
        for (int i = 0; i < 2; ++i) {
            CTRB_CUDA(cudaEventCreate(&c->kev[k][i]));
            CTRB_CUDA(cudaEventRecord(c->kev[k][i], c->stream));
        }
    CTRB_CUDA(cudaMalloc(&c->d_counters, 32 * sizeof(unsigned long long)));
    CTRB_CUDA(cudaMemset(c->d_counters, 0, 32 * sizeof(unsigned long long)));
    cudaDeviceProp prop;
    CTRB_CUDA(cudaGetDeviceProperties(&prop, device));
    c->sm_count = prop.multiProcessorCount;
    c->opt[CTRB_OPT_FAST_BLOCK_INVERT] = 1;
    c->opt[CTRB_OPT_EVAL_GROUP] = 1;
    c->opt[CTRB_OPT_GCURVE_BLOCKS] = 3;
    if (prop.major < 10) {
        set_error("device %d is sm_%d%d; libctrb200 is built for sm_100a only", device, prop.major, prop.minor);
        delete c;
        return CTRB_ERR_NO_DEVICE;
    }
    *out = c;
    return CTRB_OK;
}

void ctrb_ctx_destroy(ctrb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    cudaStreamSynchronize(c->stream2);
    for (int i = 0; i < 32; ++i)
        if (c->scratch[i]) cudaFree(c->scratch[i]);
    cudaFree(c->d_counters);
    for (int i = 0; i < 2; ++i) {
        cudaEventDestroy(c->ev[i]);
        cudaEventDestroy(c->ev_chunk[i]);
        cudaEventDestroy(c->ev_done[i]);
    }
    for (int k = 0; k < 8; ++k)
        for (int i = 0; i < 2; ++i) cudaEventDestroy(c->kev[k][i]);
    cudaStreamDestroy(c->own_stream);
    cudaStreamDestroy(c->stream2);
    delete c;
}

int ctrb_ctx_synchronize(ctrb_ctx* c) {
    CTRB_REQUIRE(c, "ctx is NULL");
    CTRB_CUDA(cudaSetDevice(c->device));
    CTRB_CUDA(cudaStreamSynchronize(c->stream));
    CTRB_CUDA(cudaStreamSynchronize(c->stream2));
    return CTRB_OK;
}

int ctrb_ctx_set_stream(ctrb_ctx* c, void* cuda_stream) {
    CTRB_REQUIRE(c, "ctx is NULL");
    CTRB_CUDA(cudaSetDevice(c->device));
    CTRB_CUDA(cudaStreamSynchronize(c->stream));
    c->stream = cuda_stream ? (cudaStream_t)cuda_stream : c->own_stream;
    return CTRB_OK;
}

int ctrb_ctx_set_option(ctrb_ctx* c, int option, int value) {
    CTRB_REQUIRE(c, "ctx is NULL");
    CTRB_REQUIRE(option >= 0 && option < CTRB_OPT_COUNT, "unknown option %d", option);
    if (option == CTRB_OPT_GCURVE_BLOCKS) CTRB_REQUIRE(value >= 1 && value <= 32, "CTRB_OPT_GCURVE_BLOCKS must be in [1, 32]");
    else if (option == CTRB_OPT_MAX_NODES_HINT) CTRB_REQUIRE(value >= 0 && value <= 1 << 20, "CTRB_OPT_MAX_NODES_HINT must be in [0, 2^20]");
    else CTRB_REQUIRE(value == 0 || value == 1, "option %d takes 0 or 1", option);
    c->opt[option] = value;
    return CTRB_OK;
}

int ctrb_ctx_get_option(const ctrb_ctx* c, int option, int* value) {
    CTRB_REQUIRE(c && value, "bad argument");
    CTRB_REQUIRE(option >= 0 && option < CTRB_OPT_COUNT, "unknown option %d", option);
    *value = c->opt[option];
    return CTRB_OK;
}

int ctrb_measure_fma_peak(ctrb_ctx* c, int precision, double* tflops) {
    CTRB_REQUIRE(c && tflops, "bad argument");
    CTRB_REQUIRE(precision == CTRB_F64 || precision == CTRB_F32, "precision must be 64 or 32");
    CTRB_CUDA(cudaSetDevice(c->device));
    return launch_fma_peak(c, precision, tflops);
}

int64_t ctrb_ctx_launch_count(const ctrb_ctx* c) { return c ? c->launches : 0; }

int ctrb_ctx_mark(ctrb_ctx* c, int which) {
    CTRB_REQUIRE(c && (which == 0 || which == 1), "ctrb_ctx_mark: bad argument");
    CTRB_CUDA(cudaSetDevice(c->device));
    CTRB_CUDA(cudaEventRecord(c->ev[which], c->stream));
    return CTRB_OK;
}

int ctrb_ctx_elapsed_ms(ctrb_ctx* c, float* ms) {
    CTRB_REQUIRE(c && ms, "ctrb_ctx_elapsed_ms: bad argument");
    CTRB_CUDA(cudaSetDevice(c->device));
    CTRB_CUDA(cudaEventSynchronize(c->ev[1]));
    CTRB_CUDA(cudaEventElapsedTime(ms, c->ev[0], c->ev[1]));
    return CTRB_OK;
}

int ctrb_ctx_last_kernel_ms(ctrb_ctx* c, int kernel_id, float* ms) {
    CTRB_REQUIRE(c && ms && kernel_id >= 0 && kernel_id < 7, "ctrb_ctx_last_kernel_ms: bad argument");
    CTRB_CUDA(cudaSetDevice(c->device));
    CTRB_CUDA(cudaEventSynchronize(c->kev[kernel_id][1]));
    CTRB_CUDA(cudaEventElapsedTime(ms, c->kev[kernel_id][0], c->kev[kernel_id][1]));
    return CTRB_OK;
}

int ctrb_static_last_fallbacks(ctrb_ctx* c, uint64_t* count) {
    CTRB_REQUIRE(c && count, "ctrb_static_last_fallbacks: bad argument");
    CTRB_CUDA(cudaSetDevice(c->device));
    CTRB_CUDA(cudaStreamSynchronize(c->stream));
    unsigned long long h = 0;
    CTRB_CUDA(cudaMemcpy(&h, c->d_counters + 11, sizeof h, cudaMemcpyDeviceToHost));
    *count = h;
    return CTRB_OK;
}

int ctrb_static_last_fallback_reasons(ctrb_ctx* c, uint64_t out[6]) {
    CTRB_REQUIRE(c && out, "ctrb_static_last_fallback_reasons: bad argument");
    CTRB_CUDA(cudaSetDevice(c->device));
    CTRB_CUDA(cudaStreamSynchronize(c->stream));
    unsigned long long h[6];
    CTRB_CUDA(cudaMemcpy(h, c->d_counters + 16, sizeof h, cudaMemcpyDeviceToHost));
    for (int k = 0; k < 6; ++k) out[k] = h[k];
    return CTRB_OK;
}

int ctrb_ctx_device_status(ctrb_ctx* c, int32_t* flags) {
    CTRB_REQUIRE(c && flags, "ctrb_ctx_device_status: bad argument");
    CTRB_CUDA(cudaSetDevice(c->device));
    CTRB_CUDA(cudaStreamSynchronize(c->stream));
    unsigned long long h = 0;
    CTRB_CUDA(cudaMemcpy(&h, c->d_counters + 12, sizeof h, cudaMemcpyDeviceToHost));
    *flags = (int32_t)(h & 0xffffffffu);
    return CTRB_OK;
}

int ctrb_ctx_work_counters(ctrb_ctx* c, uint64_t out[4]) {
    CTRB_REQUIRE(c && out, "ctrb_ctx_work_counters: bad argument");
    CTRB_CUDA(cudaSetDevice(c->device));
    CTRB_CUDA(cudaStreamSynchronize(c->stream));
    unsigned long long h[8];
    CTRB_CUDA(cudaMemcpy(h, c->d_counters, sizeof h, cudaMemcpyDeviceToHost));
    for (int i = 0; i < 4; ++i) out[i] = h[i];
    return CTRB_OK;
}

// ------------------------------------------------------------------ model
static int check_dof(int kind, int dof) {  // strain_bases.py:316-351
    switch (kind) {
        case CTRB_BASE_CONSTANT: return dof == 1;
        case CTRB_BASE_LINEAR:
        case CTRB_BASE_QUADRATIC: return dof >= 2 && dof <= 3;
        case CTRB_BASE_PURE_HELIX:
        case CTRB_BASE_LINEAR_HELIX: return dof == 2;
        case CTRB_BASE_TORSION_HELIX:
        case CTRB_BASE_TORSION_LINEAR_HELIX: return dof == 3;
        case CTRB_BASE_FULL: return dof >= 5 && dof <= 8;
        default: return 0;
    }
}

int ctrb_model_create(ctrb_ctx* ctx, const ctrb_model_desc* d, ctrb_model** out) {
    CTRB_REQUIRE(ctx && d && out, "ctrb_model_create: NULL argument");
    *out = nullptr;
    CTRB_REQUIRE(d->tube_num >= 1 && d->tube_num <= CTRB_MAX_TUBES, "tube_num %d not in 1..%d", d->tube_num, CTRB_MAX_TUBES);
    CTRB_REQUIRE(d->delta_x > 0.0, "delta_x must be positive");
    CTRB_REQUIRE(d->model_type == CTRB_MODEL_KINEMATIC || d->model_type == CTRB_MODEL_STATIC, "unknown model_type %d",
                 d->model_type);
    ctrb_model* m = new ctrb_model();
    std::memset(m, 0, sizeof *m);
    m->desc = *d;
    m->device = ctx->device;
    ModelConst& mc = m->mc;
    mc.T = d->tube_num;
    mc.dx = d->delta_x;
    int tot = 0;
    for (int n = 0; n < mc.T; ++n) {
        if (!check_dof(d->base_kind[n], d->q_dof[n])) {
            set_error("strain base %d should not have %d degrees of freedom", d->base_kind[n], d->q_dof[n]);
            delete m;
            return CTRB_ERR_INVALID;
        }
        mc.kind[n] = d->base_kind[n];
        mc.dof[n] = d->q_dof[n];
        for (int k = 0; k < CTRB_MAX_DOF; ++k) mc.q[n][k] = k < d->q_dof[n] ? d->q[n][k] : 0.0;
        mc.L[n] = d->tube_length[n];
        mc.npts[n] = (int)std::nearbyint(d->tube_length[n] / d->delta_x) + 1;  // model.py:37-38 (round half even)
        if (mc.npts[n] < 1) {
            set_error("tube %d has no discrete points", n);
            delete m;
            return CTRB_ERR_INVALID;
        }
        mc.node_base[n] = tot;
        tot += mc.npts[n];
    }
    mc.total_nodes = tot;
    if (d->model_type == CTRB_MODEL_STATIC) {
        int rcs = static_model_init(d, mc, &m->sm);
        if (rcs) {
            delete m;
            return rcs;
        }
        m->is_static = 1;
    }
    // every failure below releases the handle and whatever it already owns (ctrb_model_destroy frees null pointers too)
    auto build = [&]() -> int {
        CTRB_CUDA(cudaSetDevice(ctx->device));
        CTRB_CUDA(cudaMalloc(&m->d_tab, (size_t)12 * tot * sizeof(double)));
        CTRB_CUDA(cudaMalloc(&m->d_tab_inv, (size_t)12 * tot * sizeof(double)));
        CTRB_CUDA(cudaMalloc(&m->d_tabR32, (size_t)9 * tot * sizeof(float)));
        int rc = launch_table(ctx, m);
        if (rc == CTRB_OK && m->is_static) rc = launch_static_tables(ctx, m);
        if (rc) return rc;
        CTRB_CUDA(cudaStreamSynchronize(ctx->stream));
        return CTRB_OK;
    };
    const int rc = build();
    if (rc) {
        ctrb_model_destroy(m);
        return rc;
    }
    *out = m;
    return CTRB_OK;
}

void ctrb_model_destroy(ctrb_model* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    cudaFree(m->d_tab);
    cudaFree(m->d_tab_inv);
    cudaFree(m->d_tabR32);
    cudaFree(m->d_ws_tab);
    cudaFree(m->d_b3_tab);
    cudaFree(m->d_quad_tab);
    cudaFree(m->d_c33_tab);
    delete m;
}

int ctrb_model_num_points(const ctrb_model* m, int32_t* npts) {
    CTRB_REQUIRE(m && npts, "ctrb_model_num_points: NULL argument");
    for (int n = 0; n < m->mc.T; ++n) npts[n] = m->mc.npts[n];
    return CTRB_OK;
}

// ------------------------------------------------------------------ env
static int upload_bvh(const HostBvh& hb, DevBvh* db, void** d_nodes, void** d_tri32, void** d_tri64) {
    std::memset(db, 0, sizeof *db);
    *d_nodes = *d_tri32 = *d_tri64 = nullptr;
    if (hb.ntri <= 0) return CTRB_OK;
    CTRB_CUDA(cudaMalloc(d_nodes, hb.nodes.size() * sizeof(BvhNode)));
    CTRB_CUDA(cudaMalloc(d_tri32, hb.tri32.size() * sizeof(float)));
    CTRB_CUDA(cudaMalloc(d_tri64, hb.tri64.size() * sizeof(double)));
    CTRB_CUDA(cudaMemcpy(*d_nodes, hb.nodes.data(), hb.nodes.size() * sizeof(BvhNode), cudaMemcpyHostToDevice));
    CTRB_CUDA(cudaMemcpy(*d_tri32, hb.tri32.data(), hb.tri32.size() * sizeof(float), cudaMemcpyHostToDevice));
    CTRB_CUDA(cudaMemcpy(*d_tri64, hb.tri64.data(), hb.tri64.size() * sizeof(double), cudaMemcpyHostToDevice));
    db->nodes = (const BvhNode*)*d_nodes;
    db->tri32 = (const float4*)*d_tri32;
    db->tri64 = (const double*)*d_tri64;
    db->nnodes = (int)hb.nodes.size();
    db->ntri = hb.ntri;
    db->slack = hb.slack;
    return CTRB_OK;
}

static void copy_prim(const ctrb_prim& s, DevPrim* d) {
    d->shape = s.shape;
    for (int k = 0; k < 3; ++k) {
        d->center[k] = s.center[k];
        d->dims[k] = s.dims[k];
    }
    for (int k = 0; k < 9; ++k) d->rot[k] = s.rot[k];
}

int ctrb_env_create(ctrb_ctx* ctx, const double* obstacle_tris, int32_t n_obs, const ctrb_prim* prims, int32_t n_prims,
                    const ctrb_prim* goal, const double* goal_tris, int32_t n_goal, ctrb_env** out) {
    CTRB_REQUIRE(ctx && out, "ctrb_env_create: NULL argument");
    *out = nullptr;
    CTRB_REQUIRE(n_obs >= 0 && n_goal >= 0 && n_prims >= 0, "negative count");
    CTRB_REQUIRE(n_obs == 0 || obstacle_tris, "obstacle_tris is NULL");
    CTRB_REQUIRE(n_prims == 0 || prims, "prims is NULL");
    CTRB_REQUIRE(n_prims <= kMaxPrims, "at most %d primitive obstacles are supported", kMaxPrims);
    for (int i = 0; i < n_prims; ++i) {
        if (prims[i].shape != CTRB_SHAPE_SPHERE && prims[i].shape != CTRB_SHAPE_BOX &&
            prims[i].shape != CTRB_SHAPE_CYLINDER) {
            set_error("Shape type %d not supported.", prims[i].shape);  // init_collision.py:140
            return CTRB_ERR_UNSUPPORTED;
        }
    }
    if (goal && goal->shape == CTRB_SHAPE_MESH) CTRB_REQUIRE(n_goal > 0 && goal_tris, "mesh goal without triangles");
    CTRB_CUDA(cudaSetDevice(ctx->device));
    ctrb_env* e = new ctrb_env();
    std::memset(e, 0, sizeof *e);
    e->device = ctx->device;
    HostBvh hb, hg;
    build_bvh(obstacle_tris, n_obs, &hb);
    build_bvh(goal_tris, goal && goal->shape == CTRB_SHAPE_MESH ? n_goal : 0, &hg);
    if (hb.depth > 44 || hg.depth > 44) {
        set_error("BVH too deep (%d) for the traversal stack", std::max(hb.depth, hg.depth));
        delete e;
        return CTRB_ERR_UNSUPPORTED;
    }
    int rc = upload_bvh(hb, &e->ec.obs, &e->d_nodes[0], &e->d_tri32[0], &e->d_tri64[0]);
    if (!rc) rc = upload_bvh(hg, &e->ec.goal_mesh, &e->d_nodes[1], &e->d_tri32[1], &e->d_tri64[1]);
    if (rc) {
        ctrb_env_destroy(e);
        return rc;
    }
    e->ec.nprim = n_prims;
    for (int i = 0; i < n_prims; ++i) copy_prim(prims[i], &e->ec.prims[i]);
    if (goal) copy_prim(*goal, &e->ec.goal);
    else e->ec.goal.shape = CTRB_SHAPE_NONE;
    CTRB_CUDA(cudaMalloc(&e->d_ec, sizeof(EnvConst)));
    CTRB_CUDA(cudaMemcpy(e->d_ec, &e->ec, sizeof(EnvConst), cudaMemcpyHostToDevice));
    e->stats[0] = (int32_t)hb.nodes.size();
    e->stats[1] = hb.nleaves;
    e->stats[2] = hb.depth;
    e->stats[3] = hb.ntri;
    *out = e;
    return CTRB_OK;
}

void ctrb_env_destroy(ctrb_env* e) {
    if (!e) return;
    cudaSetDevice(e->device);
    for (int i = 0; i < 2; ++i) {
        cudaFree(e->d_nodes[i]);
        cudaFree(e->d_tri32[i]);
        cudaFree(e->d_tri64[i]);
    }
    cudaFree(e->d_ec);
    delete e;
}

int ctrb_env_stats(const ctrb_env* e, int32_t out[4]) {
    CTRB_REQUIRE(e && out, "ctrb_env_stats: NULL argument");
    for (int i = 0; i < 4; ++i) out[i] = e->stats[i];
    return CTRB_OK;
}

#define CTRB_ENTER(ctx)                                 \
    CTRB_REQUIRE((ctx) != nullptr, "ctx is NULL");      \
    CTRB_CUDA(cudaSetDevice((ctx)->device))

int ctrb_calculate_indices_batch(ctrb_ctx* ctx, const ctrb_model* m, int64_t B, const double* prev_ins,
                                 const double* next_ins, int32_t* prev_idx, int32_t* new_idx, int mem) {
    CTRB_ENTER(ctx);
    CTRB_REQUIRE(m && B >= 0, "bad argument");
    if (B == 0) return CTRB_OK;
    CTRB_REQUIRE(prev_ins && next_ins && prev_idx && new_idx, "NULL buffer");
    const size_t n = (size_t)B * m->mc.T;
    Stage st(ctx);
    const double* dp = st.in(prev_ins, n, mem);
    const double* dn = st.in(next_ins, n, mem);
    int32_t* op = st.out(prev_idx, n, mem);
    int32_t* on = st.out(new_idx, n, mem);
    if (st.rc) return st.rc;
    int rc = launch_calc_indices(ctx, m, B, dp, dn, op, on);
    if (rc) return rc;
    st.back(prev_idx, op, n, mem);
    st.back(new_idx, on, n, mem);
    return st.finish(mem);
}

int ctrb_g_node_offsets(ctrb_ctx* ctx, const ctrb_model* m, int64_t B, const int32_t* idx, int full,
                        int64_t* node_offsets, int mem) {
    CTRB_ENTER(ctx);
    CTRB_REQUIRE(m && B >= 0 && node_offsets, "bad argument");
    CTRB_REQUIRE(B == 0 || idx, "idx is NULL");
    const size_t n = (size_t)B * m->mc.T;
    if (mem == CTRB_MEM_HOST) {
        // validate on the host: indices outside [0, npts-1] are an error (the reference would index out of range)
        for (size_t i = 0; i < n; ++i) {
            int t = (int)(i % m->mc.T);
            CTRB_REQUIRE(idx[i] >= 0 && idx[i] < m->mc.npts[t], "idx[%zu]=%d outside [0,%d)", i, idx[i], m->mc.npts[t]);
        }
    }
    Stage st(ctx);
    const int32_t* di = st.in(idx, n, mem);
    long long* off = (long long*)st.out(node_offsets, n + 1, mem);
    if (st.rc) return st.rc;
    int rc = launch_offsets(ctx, m, B, di, full, off);
    if (rc) return rc;
    st.back(node_offsets, (int64_t*)off, n + 1, mem);
    return st.finish(mem);
}

int ctrb_solve_g_batch(ctrb_ctx* ctx, const ctrb_model* m, int64_t B, const int32_t* idx, const double* theta, int full,
                       int precision, const int64_t* node_offsets, void* g_out, int mem) {
    CTRB_ENTER(ctx);
    CTRB_REQUIRE(m && B >= 0, "bad argument");
    CTRB_REQUIRE(precision == CTRB_F64 || precision == CTRB_F32, "precision must be 64 or 32");
    if (B == 0) return CTRB_OK;
    CTRB_REQUIRE(idx && theta && node_offsets && g_out, "NULL buffer");
    const size_t n = (size_t)B * m->mc.T;
    const size_t esz = precision == CTRB_F64 ? 8 : 4;
    if (mem == CTRB_MEM_DEVICE) {
        return launch_gcurve(ctx, m, B, idx, theta, full, precision, (const long long*)node_offsets, g_out);
    }
    for (size_t i = 0; i < n; ++i) {
        int t = (int)(i % m->mc.T);
        CTRB_REQUIRE(idx[i] >= 0 && idx[i] < m->mc.npts[t], "idx[%zu]=%d outside [0,%d)", i, idx[i], m->mc.npts[t]);
    }
    const size_t total_nodes = (size_t)node_offsets[n];
    Stage st(ctx);
    const int32_t* di = st.in(idx, n, mem);
    const double* dt = st.in(theta, n, mem);
    const int64_t* doff = st.in(node_offsets, n + 1, mem);
    if (st.rc) return st.rc;
    void* dg = nullptr;
    int rc = ctx_scratch(ctx, 10, total_nodes * 16 * esz, &dg);
    if (rc) return rc;
    rc = launch_gcurve(ctx, m, B, di, dt, full, precision, (const long long*)doff, dg);
    if (rc) return rc;
    CTRB_CUDA(cudaMemcpyAsync(g_out, dg, total_nodes * 16 * esz, cudaMemcpyDeviceToHost, ctx->stream));
    return st.finish(mem);
}

int ctrb_eta_ftl_batch(ctrb_ctx* ctx, const ctrb_model* m, int64_t B, const int32_t* prev_idx, const int32_t* new_idx,
                       const double* velocity, const double* delta_theta, const double* prev_g, const int64_t* prev_offsets, double* eta_out,
                       double* ftl_out, double* ftl_mag, int mem) {
    CTRB_ENTER(ctx);
    CTRB_REQUIRE(m && B >= 0, "bad argument");
    if (B == 0) return CTRB_OK;
    CTRB_REQUIRE(prev_idx && delta_theta && prev_g && prev_offsets && eta_out, "NULL buffer");
    CTRB_REQUIRE((new_idx != nullptr) != (velocity != nullptr), "pass exactly one of new_idx and velocity");
    const size_t n = (size_t)B * m->mc.T;
    // device-side error flag of this call ("prev_g too short"): read back below for host buffers, through
    // ctrb_ctx_device_status for device buffers (no synchronisation here)
    void* dflag = ctx->d_counters + 12;
    int rc = CTRB_OK;
    CTRB_CUDA(cudaMemsetAsync(dflag, 0, sizeof(unsigned long long), ctx->stream));
    if (mem == CTRB_MEM_DEVICE) {
        return launch_eta_ftl(ctx, m, B, prev_idx, new_idx, velocity, delta_theta, prev_g, (const long long*)prev_offsets, eta_out,
                              ftl_out, ftl_mag, (int*)dflag);
    }
    const size_t total_nodes = (size_t)prev_offsets[n];
    Stage st(ctx);
    const int32_t* dpi = st.in(prev_idx, n, mem);
    const int32_t* dni = new_idx ? st.in(new_idx, n, mem) : nullptr;
    const double* dve = velocity ? st.in(velocity, n, mem) : nullptr;
    const double* ddt = st.in(delta_theta, n, mem);
    const int64_t* doff = st.in(prev_offsets, n + 1, mem);
    double* deta = st.out(eta_out, n * 6, mem);
    double* dmag = st.out(ftl_mag, (size_t)B * 4, mem);
    if (st.rc) return st.rc;
    void* dg = nullptr;
    rc = ctx_scratch(ctx, 10, total_nodes * 16 * sizeof(double), &dg);
    if (rc) return rc;
    CTRB_CUDA(cudaMemcpyAsync(dg, prev_g, total_nodes * 16 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    double* dftl = nullptr;
    if (ftl_out) {
        void* p = nullptr;
        rc = ctx_scratch(ctx, 11, total_nodes * 6 * sizeof(double), &p);
        if (rc) return rc;
        dftl = (double*)p;
    }
    rc = launch_eta_ftl(ctx, m, B, dpi, dni, dve, ddt, (const double*)dg, (const long long*)doff, deta, dftl, dmag, (int*)dflag);
    if (rc) return rc;
    st.back(eta_out, deta, n * 6, mem);
    st.back(ftl_mag, dmag, (size_t)B * 4, mem);
    if (ftl_out) CTRB_CUDA(cudaMemcpyAsync(ftl_out, dftl, total_nodes * 6 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    int hflag = 0;
    CTRB_CUDA(cudaMemcpyAsync(&hflag, dflag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    rc = st.finish(mem);
    if (rc) return rc;
    CTRB_REQUIRE(hflag == 0, "prev_g holds fewer nodes than npts - prev_idx for some tube (IndexError in the reference)");
    return CTRB_OK;
}

int ctrb_check_collision_batch(ctrb_ctx* ctx, const ctrb_env* env, int64_t B, int32_t tube_num, const double* g,
                               const int64_t* node_offsets, const double* rad, double* obstacle_min, double* goal_dist,
                               int mem) {
    CTRB_ENTER(ctx);
    CTRB_REQUIRE(env && B >= 0, "bad argument");
    CTRB_REQUIRE(tube_num >= 1 && tube_num <= CTRB_MAX_TUBES, "tube_num out of range");
    if (B == 0) return CTRB_OK;
    CTRB_REQUIRE(g && node_offsets && rad && obstacle_min && goal_dist, "NULL buffer");
    const size_t n = (size_t)B * tube_num;
    if (mem == CTRB_MEM_DEVICE && ctx->opt[CTRB_OPT_MAX_NODES_HINT] > 0) {
        // the caller states the longest curve (nodes per configuration): nothing to read back, no synchronisation
        return launch_eval_curves(ctx, env, B, tube_num, g, (const long long*)node_offsets, ctx->opt[CTRB_OPT_MAX_NODES_HINT], rad,
                                  obstacle_min, goal_dist);
    }
    if (mem == CTRB_MEM_DEVICE) {
        // max nodes per configuration is needed for the shared-memory carve-up: without CTRB_OPT_MAX_NODES_HINT the
        // offsets are read back (one stream synchronisation, documented in the header)
        std::vector<int64_t> h(n + 1);
        CTRB_CUDA(cudaMemcpyAsync(h.data(), node_offsets, (n + 1) * sizeof(int64_t), cudaMemcpyDeviceToHost, ctx->stream));
        CTRB_CUDA(cudaStreamSynchronize(ctx->stream));
        int64_t mx = 1;
        for (int64_t b = 0; b < B; ++b) mx = std::max(mx, h[(b + 1) * tube_num] - h[b * tube_num]);
        return launch_eval_curves(ctx, env, B, tube_num, g, (const long long*)node_offsets, (int)mx, rad, obstacle_min, goal_dist);
    }
    int64_t mx = 1;
    for (int64_t b = 0; b < B; ++b) mx = std::max(mx, node_offsets[(b + 1) * tube_num] - node_offsets[b * tube_num]);
    const size_t total_nodes = (size_t)node_offsets[n];
    Stage st(ctx);
    const int64_t* doff = st.in(node_offsets, n + 1, mem);
    double* dob = st.out(obstacle_min, (size_t)B, mem);
    double* dgo = st.out(goal_dist, (size_t)B, mem);
    if (st.rc) return st.rc;
    void* dg = nullptr;
    int rc = ctx_scratch(ctx, 10, total_nodes * 16 * sizeof(double), &dg);
    if (rc) return rc;
    CTRB_CUDA(cudaMemcpyAsync(dg, g, total_nodes * 16 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    rc = launch_eval_curves(ctx, env, B, tube_num, (const double*)dg, (const long long*)doff, (int)mx, rad, dob, dgo);
    if (rc) return rc;
    st.back(obstacle_min, dob, (size_t)B, mem);
    st.back(goal_dist, dgo, (size_t)B, mem);
    return st.finish(mem);
}

int ctrb_eval_batch(ctrb_ctx* ctx, const ctrb_model* m, const ctrb_env* env, const ctrb_cost_desc* cd, int64_t B,
                    const int32_t* idx, const double* theta, const double* rad, double* obstacle_min, double* goal_dist,
                    double* cost, uint8_t* collide, int mem) {
    CTRB_ENTER(ctx);
    CTRB_REQUIRE(m && env && cd && B >= 0, "bad argument");
    if (B == 0) return CTRB_OK;
    CTRB_REQUIRE(idx && theta && rad && obstacle_min && goal_dist, "NULL buffer");
    const int T = m->mc.T;
    if (mem == CTRB_MEM_DEVICE) {
        return launch_eval_fused(ctx, m, env, cd, B, idx, theta, rad, obstacle_min, goal_dist, cost, collide);
    }
    const size_t n = (size_t)B * T;
    const int64_t CH = pipe_chunk(B);
    if (CH < B) {  // large batch: chunked copy / compute pipeline
        void* d[2][6];
        const size_t sz[6] = {(size_t)CH * T * sizeof(int32_t), (size_t)CH * T * sizeof(double), (size_t)CH * sizeof(double),
                              (size_t)CH * sizeof(double), (size_t)CH * sizeof(double), (size_t)CH};
        for (int p = 0; p < 2; ++p)
            for (int k = 0; k < 6; ++k) {
                int rc = ctx_scratch(ctx, (p ? 17 : 1) + k, sz[k], &d[p][k]);
                if (rc) return rc;
            }
        return host_pipeline(
            ctx, B, CH,
            [&](int64_t b0, int64_t nb, int p, cudaStream_t s) -> int {
                int rc = check_idx_host(m, nb, idx + b0 * T, b0);
                if (rc) return rc;
                CTRB_COPY(d[p][0], idx + b0 * T, (size_t)nb * T * sizeof(int32_t), cudaMemcpyHostToDevice, s);
                CTRB_COPY(d[p][1], theta + b0 * T, (size_t)nb * T * sizeof(double), cudaMemcpyHostToDevice, s);
                return (int)CTRB_OK;
            },
            [&](int64_t, int64_t nb, int p) -> int {
                return launch_eval_fused(ctx, m, env, cd, nb, (const int32_t*)d[p][0], (const double*)d[p][1], rad,
                                         (double*)d[p][2], (double*)d[p][3], cost ? (double*)d[p][4] : nullptr,
                                         collide ? (uint8_t*)d[p][5] : nullptr);
            },
            [&](int64_t b0, int64_t nb, int p, cudaStream_t s) -> int {
                CTRB_COPY(obstacle_min + b0, d[p][2], (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost, s);
                CTRB_COPY(goal_dist + b0, d[p][3], (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost, s);
                if (cost) CTRB_COPY(cost + b0, d[p][4], (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost, s);
                if (collide) CTRB_COPY(collide + b0, d[p][5], (size_t)nb, cudaMemcpyDeviceToHost, s);
                return (int)CTRB_OK;
            });
    }
    {
        int rc = check_idx_host(m, B, idx);
        if (rc) return rc;
    }
    Stage st(ctx);
    const int32_t* di = st.in(idx, n, mem);
    const double* dt = st.in(theta, n, mem);
    double* dob = st.out(obstacle_min, (size_t)B, mem);
    double* dgo = st.out(goal_dist, (size_t)B, mem);
    double* dco = st.out(cost, (size_t)B, mem);
    uint8_t* dcl = st.out(collide, (size_t)B, mem);
    if (st.rc) return st.rc;
    int rc = launch_eval_fused(ctx, m, env, cd, B, di, dt, rad, dob, dgo, dco, dcl);
    if (rc) return rc;
    st.back(obstacle_min, dob, (size_t)B, mem);
    st.back(goal_dist, dgo, (size_t)B, mem);
    st.back(cost, dco, (size_t)B, mem);
    st.back(collide, dcl, (size_t)B, mem);
    return st.finish(mem);
}

// ------------------------------------------------------------------ static model
int ctrb_static_num_unknowns(const ctrb_model* m, int32_t* nq) {
    CTRB_REQUIRE(m && nq, "NULL argument");
    CTRB_REQUIRE(m->is_static, "not a static model handle");
    *nq = m->sm.nq;
    return CTRB_OK;
}

int ctrb_static_initial_guess(const ctrb_model* m, double* x0) {
    CTRB_REQUIRE(m && x0, "NULL argument");
    CTRB_REQUIRE(m->is_static, "not a static model handle");
    for (int k = 0; k < m->sm.nq; ++k) x0[k] = m->sm.x0[k];
    return CTRB_OK;
}


int ctrb_static_residual_batch(ctrb_ctx* ctx, const ctrb_model* m, int64_t B, const int32_t* idx, const double* theta,
                               const double* q, double* F, int mem) {
    CTRB_ENTER(ctx);
    CTRB_REQUIRE(m && B >= 0, "bad argument");
    CTRB_REQUIRE(m->is_static, "not a static model handle");
    if (B == 0) return CTRB_OK;
    CTRB_REQUIRE(idx && theta && q && F, "NULL buffer");
    const size_t n = (size_t)B * m->mc.T, nqs = (size_t)B * m->sm.nq;
    if (mem == CTRB_MEM_HOST) {
        int rc = check_idx_host(m, B, idx);
        if (rc) return rc;
    }
    Stage st(ctx);
    const int32_t* di = st.in(idx, n, mem);
    const double* dt = st.in(theta, n, mem);
    const double* dq = st.in(q, nqs, mem);
    double* dF = st.out(F, nqs, mem);
    if (st.rc) return st.rc;
    int rc = launch_static_residual(ctx, m->sm, B, di, dt, dq, dF);
    if (rc) return rc;
    st.back(F, dF, nqs, mem);
    return st.finish(mem);
}

int ctrb_static_solve_g_batch(ctrb_ctx* ctx, const ctrb_model* m, int64_t B, const int32_t* idx, const double* theta,
                              const double* initial_guess, int x0_per_config, double xtol, const int64_t* node_offsets,
                              double* g_out, double* solution_q, int32_t* nfev, int32_t* info, int mem) {
    CTRB_ENTER(ctx);
    CTRB_REQUIRE(m && B >= 0, "bad argument");
    CTRB_REQUIRE(m->is_static, "not a static model handle");
    if (B == 0) return CTRB_OK;
    CTRB_REQUIRE(idx && theta, "NULL buffer");
    CTRB_REQUIRE(!g_out || node_offsets, "g_out needs node_offsets");
    const size_t n = (size_t)B * m->mc.T, nqs = (size_t)B * m->sm.nq;
    CTRB_CUDA(cudaMemsetAsync(ctx->d_counters + 11, 0, sizeof(unsigned long long), ctx->stream));
    CTRB_CUDA(cudaMemsetAsync(ctx->d_counters + 16, 0, 6 * sizeof(unsigned long long), ctx->stream));
    if (mem == CTRB_MEM_DEVICE) {
        return launch_static_solve(ctx, m->sm, B, idx, theta, initial_guess, x0_per_config, xtol,
                                   (const long long*)node_offsets, g_out, solution_q, nfev, info, 0);
    }
    int rc = check_idx_host(m, B, idx);
    if (rc) return rc;
    Stage st(ctx);
    const int32_t* di = st.in(idx, n, mem);
    const double* dt = st.in(theta, n, mem);
    const double* dx0 = initial_guess ? st.in(initial_guess, x0_per_config ? nqs : (size_t)m->sm.nq, mem) : nullptr;
    const int64_t* doff = node_offsets ? st.in(node_offsets, n + 1, mem) : nullptr;
    double* dsol = st.out(solution_q, nqs, mem);
    int32_t* dnf = st.out(nfev, (size_t)B, mem);
    int32_t* dinf = st.out(info, (size_t)B, mem);
    if (st.rc) return st.rc;
    void* dg = nullptr;
    size_t total_nodes = 0;
    if (g_out) {
        total_nodes = (size_t)node_offsets[n];
        rc = ctx_scratch(ctx, 10, total_nodes * 16 * sizeof(double), &dg);
        if (rc) return rc;
    }
    rc = launch_static_solve(ctx, m->sm, B, di, dt, dx0, x0_per_config, xtol, (const long long*)doff, (double*)dg, dsol, dnf,
                             dinf, 0);
    if (rc) return rc;
    st.back(solution_q, dsol, nqs, mem);
    st.back(nfev, dnf, (size_t)B, mem);
    st.back(info, dinf, (size_t)B, mem);
    if (g_out) CTRB_CUDA(cudaMemcpyAsync(g_out, dg, total_nodes * 16 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return st.finish(mem);
}

int ctrb_static_curves_batch(ctrb_ctx* ctx, const ctrb_model* m, int64_t B, const int32_t* idx, const double* theta,
                             const double* q, const int64_t* node_offsets, double* g_out, int mem) {
    CTRB_ENTER(ctx);
    CTRB_REQUIRE(m && B >= 0, "bad argument");
    CTRB_REQUIRE(m->is_static, "not a static model handle");
    if (B == 0) return CTRB_OK;
    CTRB_REQUIRE(idx && theta && q && node_offsets && g_out, "NULL buffer");
    const size_t n = (size_t)B * m->mc.T, nqs = (size_t)B * m->sm.nq;
    if (mem == CTRB_MEM_DEVICE)
        return launch_static_curves(ctx, m->sm, B, idx, theta, q, (const long long*)node_offsets, g_out, 0);
    int rc = check_idx_host(m, B, idx);
    if (rc) return rc;
    Stage st(ctx);
    const int32_t* di = st.in(idx, n, mem);
    const double* dt = st.in(theta, n, mem);
    const double* dq = st.in(q, nqs, mem);
    const int64_t* doff = st.in(node_offsets, n + 1, mem);
    if (st.rc) return st.rc;
    const size_t total_nodes = (size_t)node_offsets[n];
    void* dg = nullptr;
    rc = ctx_scratch(ctx, 10, total_nodes * 16 * sizeof(double), &dg);
    if (rc) return rc;
    rc = launch_static_curves(ctx, m->sm, B, di, dt, dq, (const long long*)doff, (double*)dg, 0);
    if (rc) return rc;
    CTRB_CUDA(cudaMemcpyAsync(g_out, dg, total_nodes * 16 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    return st.finish(mem);
}

int ctrb_static_eval_batch(ctrb_ctx* ctx, const ctrb_model* m, const ctrb_env* env, const ctrb_cost_desc* cd, int64_t B,
                           const int32_t* idx, const double* theta, const double* rad, double* obstacle_min,
                           double* goal_dist, double* cost, uint8_t* collide, int32_t* info, int32_t* nfev, int mem) {
    CTRB_ENTER(ctx);
    CTRB_REQUIRE(m && env && cd && B >= 0, "bad argument");
    CTRB_REQUIRE(m->is_static, "not a static model handle");
    if (B == 0) return CTRB_OK;
    CTRB_REQUIRE(idx && theta && rad && obstacle_min && goal_dist, "NULL buffer");
    const int T = m->mc.T;
    const size_t n = (size_t)B * T;
    // The ragged offsets and the curves live on the device only.  The collision kernel reads node positions, so the curves
    // kernel writes positions (24 B per node, not 128).  Their scratch is sized for the longest possible curves
    // (every tube fully exposed), not from the batch's own node count: reading that count back would be a hidden
    // synchronisation in CTRB_MEM_DEVICE mode.  The scratch is bounded at 64 GiB of the 180: larger batches run as
    // consecutive launch sequences on the stream (C3: 201 nodes x 24 B per configuration -> up to 14.2 M configurations
    // per sequence, so BASELINE's 10^7 is one).
    CTRB_CUDA(cudaMemsetAsync(ctx->d_counters + 11, 0, sizeof(unsigned long long), ctx->stream));
    CTRB_CUDA(cudaMemsetAsync(ctx->d_counters + 16, 0, 6 * sizeof(unsigned long long), ctx->stream));
    const int64_t per_cfg = (int64_t)m->mc.total_nodes * 3 * (int64_t)sizeof(double);
    const int64_t seq_max = std::max<int64_t>(((int64_t)64 << 30) / per_cfg, 1 << 16);
    // one launch sequence (offset scan, solve, curves, collision + cost) over nb configurations in device memory
    auto run = [&](int64_t nb, const int32_t* di, const double* dt, double* dob, double* dgo, double* dco, uint8_t* dcl,
                   int32_t* dinf, int32_t* dnf) -> int {
        const int64_t CH = std::min<int64_t>(nb, seq_max);
        void* doff = nullptr;
        int rc = ctx_scratch(ctx, 9, ((size_t)CH * T + 1) * sizeof(long long), &doff);
        if (rc) return rc;
        void* dg = nullptr;
        rc = ctx_scratch(ctx, 10, (size_t)CH * m->mc.total_nodes * 3 * sizeof(double), &dg);
        if (rc) return rc;
        for (int64_t b0 = 0; b0 < nb; b0 += CH) {
            const int64_t k = std::min(CH, nb - b0);
            rc = launch_offsets(ctx, m, k, di + b0 * T, 0, (long long*)doff);
            if (rc) return rc;
            rc = launch_static_solve(ctx, m->sm, k, di + b0 * T, dt + b0 * T, nullptr, 0, 0.0, (const long long*)doff, (double*)dg,
                                     nullptr, dnf ? dnf + b0 : nullptr, dinf ? dinf + b0 : nullptr,
                                     // node positions: one thread per configuration when there are enough configurations to
                                     // fill the machine with threads, the warp-per-configuration scan otherwise (a thread walks
                                     // up to total_nodes Magnus steps serially: 0.5 ms of latency on a 1024-configuration batch)
                                     (ctx->opt[CTRB_OPT_STATIC_CURVES_SCAN] || k < 65536) ? 2 : 1);
            if (rc) return rc;
            rc = launch_eval_curves_cost(ctx, env, cd, k, T, (const double*)dg, (const long long*)doff, m->mc.total_nodes, rad,
                                         dob + b0, dgo + b0, dco ? dco + b0 : nullptr, dcl ? dcl + b0 : nullptr, 1);
            if (rc) return rc;
        }
        return (int)CTRB_OK;
    };
    if (mem == CTRB_MEM_DEVICE) return run(B, idx, theta, obstacle_min, goal_dist, cost, collide, info, nfev);
    // Host buffers, large batch: a copy / compute pipeline over a few chunks (the copies are 2.7 % of the call at 10^7
    // configurations; chunks stay at 2^21 or more, so that the kernels' tails -- ~1.5 ms per sequence -- cost less than the
    // overlap gains).  Smaller batches: one sequence (eight chunks of 2^17 measured 3.2 M instead of 4.7 M configs/s in round 1).
    if (B >= ((int64_t)1 << 22)) {
        const int64_t CH = std::min<int64_t>(std::max<int64_t>((B + 3) / 4, (int64_t)1 << 21), seq_max);
        void* d[2][8];
        const size_t sz[8] = {(size_t)CH * T * sizeof(int32_t), (size_t)CH * T * sizeof(double), (size_t)CH * sizeof(double),
                              (size_t)CH * sizeof(double), (size_t)CH * sizeof(double), (size_t)CH, (size_t)CH * sizeof(int32_t),
                              (size_t)CH * sizeof(int32_t)};
        for (int p = 0; p < 2; ++p)
            for (int k = 0; k < 8; ++k) {
                int rc = ctx_scratch(ctx, (p ? 17 : 1) + k, sz[k], &d[p][k]);
                if (rc) return rc;
            }
        return host_pipeline(
            ctx, B, CH,
            [&](int64_t b0, int64_t nb, int p, cudaStream_t s) -> int {
                int rc = check_idx_host(m, nb, idx + b0 * T, b0);
                if (rc) return rc;
                CTRB_COPY(d[p][0], idx + b0 * T, (size_t)nb * T * sizeof(int32_t), cudaMemcpyHostToDevice, s);
                CTRB_COPY(d[p][1], theta + b0 * T, (size_t)nb * T * sizeof(double), cudaMemcpyHostToDevice, s);
                return (int)CTRB_OK;
            },
            [&](int64_t, int64_t nb, int p) -> int {
                return run(nb, (const int32_t*)d[p][0], (const double*)d[p][1], (double*)d[p][2], (double*)d[p][3],
                           cost ? (double*)d[p][4] : nullptr, collide ? (uint8_t*)d[p][5] : nullptr,
                           info ? (int32_t*)d[p][6] : nullptr, nfev ? (int32_t*)d[p][7] : nullptr);
            },
            [&](int64_t b0, int64_t nb, int p, cudaStream_t s) -> int {
                CTRB_COPY(obstacle_min + b0, d[p][2], (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost, s);
                CTRB_COPY(goal_dist + b0, d[p][3], (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost, s);
                if (cost) CTRB_COPY(cost + b0, d[p][4], (size_t)nb * sizeof(double), cudaMemcpyDeviceToHost, s);
                if (collide) CTRB_COPY(collide + b0, d[p][5], (size_t)nb, cudaMemcpyDeviceToHost, s);
                if (info) CTRB_COPY(info + b0, d[p][6], (size_t)nb * sizeof(int32_t), cudaMemcpyDeviceToHost, s);
                if (nfev) CTRB_COPY(nfev + b0, d[p][7], (size_t)nb * sizeof(int32_t), cudaMemcpyDeviceToHost, s);
                return (int)CTRB_OK;
            });
    }
    {
        int rc = check_idx_host(m, B, idx);
        if (rc) return rc;
    }
    Stage st(ctx);
    const int32_t* di = st.in(idx, n, mem);
    const double* dt = st.in(theta, n, mem);
    double* dob = st.out(obstacle_min, (size_t)B, mem);
    double* dgo = st.out(goal_dist, (size_t)B, mem);
    double* dco = st.out(cost, (size_t)B, mem);
    uint8_t* dcl = st.out(collide, (size_t)B, mem);
    int32_t* dinf = st.out(info, (size_t)B, mem);
    int32_t* dnf = st.out(nfev, (size_t)B, mem);
    if (st.rc) return st.rc;
    int rc = run(B, di, dt, dob, dgo, dco, dcl, dinf, dnf);
    if (rc) return rc;
    st.back(obstacle_min, dob, (size_t)B, mem);
    st.back(goal_dist, dgo, (size_t)B, mem);
    st.back(cost, dco, (size_t)B, mem);
    st.back(collide, dcl, (size_t)B, mem);
    st.back(info, dinf, (size_t)B, mem);
    st.back(nfev, dnf, (size_t)B, mem);
    return st.finish(mem);
}

int ctrb_knn_batch(ctrb_ctx* ctx, int32_t dim, int64_t N, const double* points, int64_t Q, const double* queries,
                   const double* weight, const int32_t* topology, int32_t k, int32_t* idx_out, double* dist_out, int mem) {
    CTRB_ENTER(ctx);
    CTRB_REQUIRE(dim >= 1 && dim <= 2 * CTRB_MAX_TUBES, "dim must be in [1, %d]", 2 * CTRB_MAX_TUBES);
    CTRB_REQUIRE(k >= 1 && k <= 32, "k must be in [1, 32]");
    CTRB_REQUIRE(N >= 0 && Q >= 0 && weight && topology, "bad argument");
    if (Q == 0) return CTRB_OK;
    CTRB_REQUIRE(queries && idx_out && dist_out && (N == 0 || points), "NULL buffer");
    ctrb::KnnParams P;
    std::memset(&P, 0, sizeof P);
    P.dim = dim;
    P.k = k;
    P.N = N;
    P.Q = Q;
    for (int d = 0; d < dim; ++d) {
        CTRB_REQUIRE(topology[d] == 1 || topology[d] == 2, "topology[%d] = %d (1 = linear, 2 = circular)", d, topology[d]);
        P.w[d] = weight[d];
        P.topo[d] = topology[d];
    }
    Stage st(ctx);
    const double* dp = st.in(points, (size_t)std::max<int64_t>(N, 1) * dim, N > 0 ? mem : CTRB_MEM_DEVICE);
    const double* dq = st.in(queries, (size_t)Q * dim, mem);
    int32_t* di = st.out(idx_out, (size_t)Q * k, mem);
    double* dd = st.out(dist_out, (size_t)Q * k, mem);
    if (st.rc) return st.rc;
    int rc = ctrb::launch_knn(ctx, P, dp, dq, di, dd);
    if (rc) return rc;
    st.back(idx_out, di, (size_t)Q * k, mem);
    st.back(dist_out, dd, (size_t)Q * k, mem);
    return st.finish(mem);
}

int ctrb_argmin_cost(ctrb_ctx* ctx, int64_t B, const double* cost, int64_t index_base, double* best_cost,
                     int64_t* best_index, int mem) {
    CTRB_ENTER(ctx);
    CTRB_REQUIRE(B >= 0 && best_cost && best_index, "bad argument");
    *best_cost = INFINITY;
    *best_index = -1;
    if (B == 0) return CTRB_OK;
    CTRB_REQUIRE(cost, "cost is NULL");
    CTRB_REQUIRE(B < (1LL << 31), "at most 2^31-1 costs per call");
    Stage st(ctx);
    const double* dc = st.in(cost, (size_t)B, mem);
    if (st.rc) return st.rc;
    void* dkey = nullptr;
    int rc = ctx_scratch(ctx, 0, 2 * sizeof(unsigned long long), &dkey);
    if (rc) return rc;
    rc = launch_argmin(ctx, B, dc, (unsigned long long*)dkey);
    if (rc) return rc;
    unsigned long long h[2];
    CTRB_CUDA(cudaMemcpyAsync(h, dkey, sizeof h, cudaMemcpyDeviceToHost, ctx->stream));
    CTRB_CUDA(cudaStreamSynchronize(ctx->stream));
    if (h[1] != ~0ULL) {
        double c;
        unsigned long long bits = h[0];
        bits = (bits & 0x8000000000000000ULL) ? (bits & 0x7fffffffffffffffULL) : ~bits;  // undo the monotone map
        std::memcpy(&c, &bits, sizeof c);
        *best_cost = c;
        *best_index = index_base + (int64_t)h[1];
    }
    return CTRB_OK;
}

int ctrb_argmin_cost_device(ctrb_ctx* ctx, int64_t B, const double* cost, int64_t index_base, int64_t* key_index) {
    CTRB_ENTER(ctx);
    CTRB_REQUIRE(B >= 0 && key_index && (B == 0 || cost), "bad argument");
    CTRB_REQUIRE(B < (1LL << 31), "at most 2^31-1 costs per call");
    void* dkey = nullptr;
    int rc = ctx_scratch(ctx, 0, 2 * sizeof(unsigned long long), &dkey);
    if (rc) return rc;
    if (B > 0) {
        rc = launch_argmin(ctx, B, cost, (unsigned long long*)dkey);
        if (rc) return rc;
    } else {
        CTRB_CUDA(cudaMemsetAsync(dkey, 0xff, 2 * sizeof(unsigned long long), ctx->stream));
    }
    return launch_argmin_finish(ctx, (const unsigned long long*)dkey, index_base, (long long*)key_index);
}

// the monotone map of argmin_key_kernel (cost_key) with the sign bit flipped so that SIGNED comparison orders the costs
int64_t ctrb_cost_to_key(double cost) {
    unsigned long long b;
    std::memcpy(&b, &cost, sizeof b);
    b = (b & 0x8000000000000000ULL) ? ~b : (b | 0x8000000000000000ULL);
    return (int64_t)(b ^ 0x8000000000000000ULL);
}

double ctrb_key_to_cost(int64_t key) {
    if (key == INT64_MAX) return INFINITY;
    unsigned long long b = (unsigned long long)key ^ 0x8000000000000000ULL;
    b = (b & 0x8000000000000000ULL) ? (b & 0x7fffffffffffffffULL) : ~b;
    double c;
    std::memcpy(&c, &b, sizeof c);
    return c;
}

}  // extern "C"

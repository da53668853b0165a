// skb_api.cpp -- extern "C" hot-path entry points (device-pointer and host-pointer variants).
// See include/skmp_b200.h for the reference call site each one replaces.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>

#include "skb_internal.h"

using namespace skb;

#define CUDA_OK(expr)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (expr);                                                              \
        if (e__ != cudaSuccess)                                                                \
            return set_error(std::string(#expr) + ": " + cudaGetErrorString(e__));             \
    } while (0)

static size_t esize(int dtype) { return dtype == SKB_F64 ? 8 : 4; }

// A plan's program tables and an SDF's grids / primitive tables live on ONE device.  Launching with another device current
// would hand that device foreign pointers (an illegal address at best), so every compute entry point checks first.
namespace skb {
int check_plan_device(const skb_plan *p, const char *who) {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) return set_error(std::string(who) + ": no CUDA device (there is no CPU fallback)");
    if (dev != p->device)
        return set_error(std::string(who) + ": the plan was created on cuda:" + std::to_string(p->device) + " but cuda:" +
                         std::to_string(dev) + " is current (create one plan per device)");
    return 0;
}
int check_sdf_device(skb_sdf *s, const char *who) {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) return set_error(std::string(who) + ": no CUDA device (there is no CPU fallback)");
    std::lock_guard<std::mutex> lk(s->mu);
    if (s->device < 0) s->device = dev;
    if (dev != s->device)
        return set_error(std::string(who) + ": the SDF lives on cuda:" + std::to_string(s->device) + " but cuda:" +
                         std::to_string(dev) + " is current (create one SDF handle per device)");
    return 0;
}
}  // namespace skb

// ------------------------------------------------------------------ host pipeline scratch
struct skb_stream_ctx {
    static const int NS = 3;
    cudaStream_t st[NS] = {nullptr, nullptr, nullptr};
    void *dq[NS] = {nullptr, nullptr, nullptr};
    void *dv[NS] = {nullptr, nullptr, nullptr};
    void *dj[NS] = {nullptr, nullptr, nullptr};
    uint8_t *dvalid[NS] = {nullptr, nullptr, nullptr};
    size_t cap_q = 0, cap_v = 0, cap_j = 0, cap_valid = 0;
    // small batches (the planners' N = 1 calls): one pinned, device-mapped staging block -- the kernel reads q from it and
    // writes its results into it directly, so the call is one launch + one stream synchronisation, no copy engine
    void *pin = nullptr;
    size_t pin_cap = 0;
};

static const int64_t SMALL_N = 64;
static const size_t SMALL_BYTES = 1u << 20;

static int ensure_ctx(skb_plan *p, size_t bq, size_t bv, size_t bj, size_t bvalid) {
    if (!p->hostpipe) {
        p->hostpipe = new skb_stream_ctx();
        for (int i = 0; i < skb_stream_ctx::NS; ++i) CUDA_OK(cudaStreamCreateWithFlags(&p->hostpipe->st[i], cudaStreamNonBlocking));
    }
    skb_stream_ctx *c = p->hostpipe;
    auto grow = [&](void **arr, size_t &cap, size_t need) -> int {
        if (need <= cap) return 0;
        for (int i = 0; i < skb_stream_ctx::NS; ++i) {
            if (arr[i]) CUDA_OK(cudaFree(arr[i]));
            arr[i] = nullptr;
            CUDA_OK(cudaMalloc(&arr[i], need));
        }
        cap = need;
        return 0;
    };
    // every host call returns with its streams drained (also on failure) and holds p->host_mu, so nothing is in flight here
    if (grow(c->dq, c->cap_q, bq)) return -1;
    if (grow(c->dv, c->cap_v, bv)) return -1;
    if (grow(c->dj, c->cap_j, bj)) return -1;
    if (grow(reinterpret_cast<void **>(c->dvalid), c->cap_valid, bvalid)) return -1;
    return 0;
}

extern "C" void skb_plan_destroy_hostpipe(skb_plan *p) {
    if (!p || !p->hostpipe) return;
    skb_stream_ctx *c = p->hostpipe;
    if (c->pin) cudaFreeHost(c->pin);
    for (int i = 0; i < skb_stream_ctx::NS; ++i) {
        if (c->st[i]) { cudaStreamSynchronize(c->st[i]); cudaStreamDestroy(c->st[i]); }
        if (c->dq[i]) cudaFree(c->dq[i]);
        if (c->dv[i]) cudaFree(c->dv[i]);
        if (c->dj[i]) cudaFree(c->dj[i]);
        if (c->dvalid[i]) cudaFree(c->dvalid[i]);
    }
    delete c;
    p->hostpipe = nullptr;
}

// generic chunked H2D -> launch -> D2H driver.  `launch(q_dev, n, vals_dev, jac_dev, valid_dev, stream)`
template <typename Launch>
static int host_pipeline(skb_plan *p, int dtype, const void *q_host, int64_t N, size_t vals_per_cfg, size_t jac_per_cfg,
                         void *out_vals, void *out_jac, uint8_t *out_valid, Launch launch) {
    if (N <= 0) return 0;
    // one host call at a time per plan: the streams and the device scratch below belong to the plan (ctypes releases the
    // GIL, so two Python threads can be here at once)
    std::lock_guard<std::mutex> host_lock(p->host_mu);
    const size_t es = esize(dtype);
    const size_t per_cfg = (out_vals ? vals_per_cfg : 0) + (out_jac ? jac_per_cfg : 0) + (size_t)p->D;
    // ---- small batch: zero-copy through one mapped, pinned block (see skb_stream_ctx)
    static const bool zero_copy = !(getenv("SKB_HOST_ZEROCOPY") && atoi(getenv("SKB_HOST_ZEROCOPY")) == 0);
    auto up16 = [](size_t b) { return (b + 15) / 16 * 16; };
    const size_t bq = up16(N * p->D * es), bv = out_vals ? up16(N * vals_per_cfg * es) : 0, bj = out_jac ? up16(N * jac_per_cfg * es) : 0,
                 bvalid = out_valid ? up16((size_t)N) : 0;
    if (zero_copy && N <= SMALL_N && bq + bv + bj + bvalid <= SMALL_BYTES) {
        if (ensure_ctx(p, 0, 0, 0, 0)) return -1;
        skb_stream_ctx *c = p->hostpipe;
        if (c->pin_cap < SMALL_BYTES) {
            CUDA_OK(cudaHostAlloc(&c->pin, SMALL_BYTES, cudaHostAllocMapped | cudaHostAllocPortable));
            c->pin_cap = SMALL_BYTES;
        }
        char *hp = reinterpret_cast<char *>(c->pin), *dp = nullptr;
        CUDA_OK(cudaHostGetDevicePointer(reinterpret_cast<void **>(&dp), c->pin, 0));
        memcpy(hp, q_host, N * p->D * es);
        const size_t ov = bq, oj = bq + bv, ovalid = bq + bv + bj;
        int rc = launch(dp, N, out_vals ? dp + ov : nullptr, out_jac ? dp + oj : nullptr,
                        out_valid ? reinterpret_cast<uint8_t *>(dp + ovalid) : nullptr, c->st[0]);
        if (cudaStreamSynchronize(c->st[0]) != cudaSuccess && rc == 0) rc = set_error("host call: the launch failed");
        if (rc != 0) return rc;
        if (out_vals) memcpy(out_vals, hp + ov, N * vals_per_cfg * es);
        if (out_jac) memcpy(out_jac, hp + oj, N * jac_per_cfg * es);
        if (out_valid) memcpy(out_valid, hp + ovalid, (size_t)N);
        return 0;
    }
    int64_t chunk = (int64_t)std::max<size_t>(1024, (48u << 20) / std::max<size_t>(1, per_cfg * es));
    chunk = std::min<int64_t>(chunk, N);
    chunk = (chunk + 31) / 32 * 32;
    if (ensure_ctx(p, chunk * p->D * es, out_vals ? chunk * vals_per_cfg * es : 0, out_jac ? chunk * jac_per_cfg * es : 0,
                   out_valid ? (size_t)chunk : 0))
        return -1;
    skb_stream_ctx *c = p->hostpipe;
    const int used = (int)std::min<int64_t>(skb_stream_ctx::NS, (N + chunk - 1) / chunk);     // a small batch touched one stream
    auto enqueue = [&]() -> int {
        int k = 0;
        for (int64_t n0 = 0; n0 < N; n0 += chunk, k = (k + 1) % skb_stream_ctx::NS) {
            const int64_t n = std::min<int64_t>(chunk, N - n0);
            cudaStream_t st = c->st[k];
            CUDA_OK(cudaMemcpyAsync(c->dq[k], (const char *)q_host + n0 * p->D * es, n * p->D * es, cudaMemcpyHostToDevice, st));
            if (launch(c->dq[k], n, out_vals ? c->dv[k] : nullptr, out_jac ? c->dj[k] : nullptr, out_valid ? c->dvalid[k] : nullptr, st))
                return -1;
            if (out_vals)
                CUDA_OK(cudaMemcpyAsync((char *)out_vals + n0 * vals_per_cfg * es, c->dv[k], n * vals_per_cfg * es, cudaMemcpyDeviceToHost, st));
            if (out_jac)
                CUDA_OK(cudaMemcpyAsync((char *)out_jac + n0 * jac_per_cfg * es, c->dj[k], n * jac_per_cfg * es, cudaMemcpyDeviceToHost, st));
            if (out_valid) CUDA_OK(cudaMemcpyAsync(out_valid + n0, c->dvalid[k], n, cudaMemcpyDeviceToHost, st));
        }
        return 0;
    };
    const int rc = enqueue();
    // drained on success AND on failure: the caller owns its buffers again when this returns, and the next call may regrow
    // the scratch
    int sync_rc = 0;
    for (int i = 0; i < used; ++i)
        if (cudaStreamSynchronize(c->st[i]) != cudaSuccess && rc == 0 && sync_rc == 0)
            sync_rc = set_error("host pipeline: stream synchronisation failed");
    return rc != 0 ? rc : sync_rc;
}

extern "C" {

int skb_sdf_eval(const skb_sdf *s, int32_t dtype, const void *pts_dev, int64_t M, void *out_vals_dev, void *out_grads_dev,
                 void *stream) {
    if (!s || !pts_dev || !out_vals_dev) return set_error("sdf_eval: bad arguments");
    if (s->prims.empty()) return set_error("sdf_eval: the SDF has no primitive");
    if (check_sdf_device(const_cast<skb_sdf *>(s), "sdf_eval") != 0) return -1;
    const void *dp = get_dev_prims(const_cast<skb_sdf *>(s), dtype);
    if (!dp) return -1;
    return launch_sdf_eval(dp, (int)s->prims.size(), dtype, pts_dev, M, out_vals_dev, out_grads_dev, stream);
}

int skb_fk(const skb_plan *p, int32_t dtype, const void *q_dev, int64_t N, int32_t with_jacobian, void *out_pts_dev,
           void *out_jac_dev, void *stream) {
    if (!p || N < 0 || (N > 0 && (!q_dev || !out_pts_dev))) return set_error("fk: bad arguments");
    if (N > 0 && with_jacobian && !out_jac_dev) return set_error("fk: with_jacobian set but out_jac is NULL");
    if (check_plan_device(p, "fk") != 0) return -1;
    if (p->rot_type == SKB_ROT_IGNORE) {   // point features: lane-per-feature kernel (skb_sphere.cu)
        const skb_sphere_program *sp = get_sphere_program(const_cast<skb_plan *>(p));
        if (sp) return launch_sphere(p, sp, KIND_MAP, dtype, q_dev, N, nullptr, nullptr, 0, 0.0, SKB_MODE_ALL, out_pts_dev,
                                     with_jacobian ? out_jac_dev : nullptr, nullptr, stream);
    }
    else if (const skb_pose_program *pp = get_pose_program(const_cast<skb_plan *>(p)))   // RPY / XYZW rows (skb_pose.cu)
        return launch_pose(p, pp, dtype, q_dev, N, out_pts_dev, with_jacobian ? out_jac_dev : nullptr, stream);
    const skb_schedule *sch = get_schedule(const_cast<skb_plan *>(p), p->T);
    if (!sch) return -1;
    return launch_tree(p, sch, KIND_MAP, dtype, q_dev, N, nullptr, 0, 0.0, SKB_MODE_ALL, out_pts_dev,
                       with_jacobian ? out_jac_dev : nullptr, nullptr, stream);
}

int skb_collfree(const skb_plan *p, const skb_sdf *s, int32_t dtype, const void *q_dev, int64_t N, double margin,
                 int32_t mode, void *out_vals_dev, void *out_jac_dev, uint8_t *out_valid_dev, void *stream) {
    if (!p || !s || N < 0 || (N > 0 && !q_dev)) return set_error("collfree: bad arguments");
    if (mode != SKB_MODE_ALL && mode != SKB_MODE_CLOSEST) return set_error("collfree: bad mode");
    if (s->prims.empty()) return set_error("collfree: the SDF has no primitive");
    if (p->rot_type != SKB_ROT_IGNORE) return set_error("collfree: the plan must be built with rot_type IGNORE");
    if (s->prims.size() > 64) return set_error("collfree: more than 64 SDF primitives");
    if (check_plan_device(p, "collfree") != 0 || check_sdf_device(const_cast<skb_sdf *>(s), "collfree") != 0) return -1;
    const void *dp = get_dev_prims(const_cast<skb_sdf *>(s), dtype);
    if (!dp) return -1;
    // full-output fp32 mode: the two-rows-per-lane step kernel (skb_sphere2.cu)
    // the step kernel (skb_sphere2.cu) serves every mode in fp32 and fp64
    if (mode == SKB_MODE_CLOSEST || !out_valid_dev || (!out_vals_dev && !out_jac_dev))
        if (const skb_s2_program *s2 = get_s2_program(const_cast<skb_plan *>(p), KIND_COLLFREE))
            return launch_s2(p, s2, KIND_COLLFREE, dtype, q_dev, N, dp, get_host_prims(s, dtype), (int)s->prims.size(), margin, mode,
                             out_vals_dev, out_jac_dev, out_valid_dev, stream);
    if (const skb_sphere_program *sp = get_sphere_program(const_cast<skb_plan *>(p)))
        return launch_sphere(p, sp, KIND_COLLFREE, dtype, q_dev, N, dp, get_host_prims(s, dtype), (int)s->prims.size(), margin, mode, out_vals_dev,
                             out_jac_dev, out_valid_dev, stream);
    const skb_schedule *sch = get_schedule(const_cast<skb_plan *>(p), 1);
    if (!sch) return -1;
    return launch_tree(p, sch, KIND_COLLFREE, dtype, q_dev, N, dp, (int)s->prims.size(), margin, mode, out_vals_dev,
                       out_jac_dev, out_valid_dev, stream);
}

int skb_collfree_csc(const skb_plan *p, const skb_sdf *s, int32_t dtype, const void *q_dev, int64_t N, double margin,
                     void *out_vals_dev, void *out_data_dev, void *stream) {
    if (!p || !s || N < 0 || (N > 0 && (!q_dev || !out_data_dev))) return set_error("collfree_csc: bad arguments");
    if (s->prims.empty()) return set_error("collfree_csc: the SDF has no primitive");
    if (p->rot_type != SKB_ROT_IGNORE) return set_error("collfree_csc: the plan must be built with rot_type IGNORE");
    if (s->prims.size() > 64) return set_error("collfree_csc: more than 64 SDF primitives");
    if (p->D > 32) return set_error("collfree_csc: more than 32 columns (transpose the (N,S,D) blocks of skb_collfree instead)");
    if (check_plan_device(p, "collfree_csc") != 0 || check_sdf_device(const_cast<skb_sdf *>(s), "collfree_csc") != 0) return -1;
    const void *dp = get_dev_prims(const_cast<skb_sdf *>(s), dtype);
    if (!dp) return -1;
    const skb_s2_program *s2 = get_s2_program(const_cast<skb_plan *>(p), KIND_COLLFREE);
    if (!s2) return set_error("collfree_csc: this plan is not served by the step kernel (D > 64, > 255 nodes, or SKB_KERNEL set)");
    return launch_s2(p, s2, KIND_COLLFREE, dtype, q_dev, N, dp, get_host_prims(s, dtype), (int)s->prims.size(), margin, S2_MODE_ALL_CSC,
                     out_vals_dev, out_data_dev, nullptr, stream);
}

int skb_pairwise(const skb_plan *p, int32_t dtype, const void *q_dev, int64_t N, int32_t mode, void *out_vals_dev,
                 void *out_jac_dev, uint8_t *out_valid_dev, void *stream) {
    if (!p || N < 0 || (N > 0 && !q_dev)) return set_error("pairwise: bad arguments");
    if (mode != SKB_MODE_ALL && mode != SKB_MODE_CLOSEST) return set_error("pairwise: bad mode");
    if (!p->pairs_set) return set_error("pairwise: skb_plan_set_pairs was not called");
    if (check_plan_device(p, "pairwise") != 0) return -1;
    if (mode == SKB_MODE_CLOSEST || !out_valid_dev || (!out_vals_dev && !out_jac_dev))
        if (const skb_s2_program *s2 = get_s2_program(const_cast<skb_plan *>(p), KIND_PAIRWISE))
            return launch_s2(p, s2, KIND_PAIRWISE, dtype, q_dev, N, nullptr, nullptr, 0, 0.0, mode, out_vals_dev, out_jac_dev,
                             out_valid_dev, stream);
    const skb_pair_program *pp = get_pair_program(const_cast<skb_plan *>(p));
    if (!pp) return -1;
    return launch_pairwise(p, pp, dtype, q_dev, N, mode, out_vals_dev, out_jac_dev, out_valid_dev, stream);
}

int skb_com(const skb_plan *p, const skb_sdf *s, int32_t dtype, const void *q_dev, int64_t N, int32_t what, void *out_vals_dev,
            void *out_jac_dev, void *stream) {
    if (!p || N < 0 || (N > 0 && (!q_dev || !out_vals_dev))) return set_error("com: bad arguments");
    if (what != SKB_COM_POINT && what != SKB_COM_STABILITY) return set_error("com: bad `what`");
    const void *dp = nullptr;
    int n_prims = 0;
    if (check_plan_device(p, "com") != 0) return -1;
    if (what == SKB_COM_STABILITY) {
        if (!s || s->prims.empty()) return set_error("com: the stability constraint needs an SDF");
        if (s->prims.size() > 64) return set_error("com: more than 64 SDF primitives");
        if (check_sdf_device(const_cast<skb_sdf *>(s), "com") != 0) return -1;
        dp = get_dev_prims(const_cast<skb_sdf *>(s), dtype);
        if (!dp) return -1;
        n_prims = (int)s->prims.size();
    }
    const skb_com_program *cp = get_com_program(const_cast<skb_plan *>(p));
    if (!cp) return -1;
    return launch_com(p, cp, dtype, q_dev, N, dp, n_prims, what, out_vals_dev, out_jac_dev, stream);
}

int skb_plan_tune(const skb_plan *p, const skb_sdf *s, int32_t dtype, const void *q_dev, int64_t N, double margin, int32_t mode,
                  int32_t outputs, void *stream, int32_t *out_choice) {
    if (!p || !q_dev || N <= 0) return set_error("plan_tune: bad arguments");
    if (dtype != SKB_F32 && dtype != SKB_F64) return set_error("plan_tune: bad dtype");
    if ((outputs & 7) == 0) return set_error("plan_tune: no output selected (1 values, 2 jacobians, 4 validity bytes)");
    if ((outputs & 8) && (!s || !(outputs & 2) || (outputs & 4) || mode != SKB_MODE_ALL))
        return set_error("plan_tune: the transposed (CSC) layout is the all-spheres CollFreeConst Jacobian (outputs 8 | 2 [| 1], an SDF)");
    if (!s && !p->pairs_set) return set_error("plan_tune: pass an SDF (CollFreeConst) or set the plan's pairs (PairWiseSelfCollFreeConst)");
    if (check_plan_device(p, "plan_tune") != 0) return -1;
    const int64_t n = std::min<int64_t>(N, 262144);
    const size_t es = esize(dtype);
    const size_t M = mode == SKB_MODE_CLOSEST ? 1 : (s ? (size_t)p->F : p->pair_thr.size());
    void *dv = nullptr, *dj = nullptr;
    uint8_t *dvalid = nullptr;
    int rc = 0;
    if ((outputs & 1) && cudaMalloc(&dv, n * M * es) != cudaSuccess) rc = set_error("plan_tune: out of device memory");
    if (rc == 0 && (outputs & 2) && cudaMalloc(&dj, n * M * p->D * es) != cudaSuccess) rc = set_error("plan_tune: out of device memory");
    if (rc == 0 && (outputs & 4) && cudaMalloc(reinterpret_cast<void **>(&dvalid), n) != cudaSuccess) rc = set_error("plan_tune: out of device memory");
    int made = 0;
    if (rc == 0) {
        s2_tune_begin(0);
        rc = (outputs & 8) ? skb_collfree_csc(p, s, dtype, q_dev, n, margin, dv, dj, stream)
             : s           ? skb_collfree(p, s, dtype, q_dev, n, margin, mode, dv, dj, dvalid, stream)
                           : skb_pairwise(p, dtype, q_dev, n, mode, dv, dj, dvalid, stream);
        made = s2_tune_end(out_choice);
        if (cudaStreamSynchronize(reinterpret_cast<cudaStream_t>(stream)) != cudaSuccess && rc == 0) rc = set_error("plan_tune: the timed launches failed");
    }
    if (dv) cudaFree(dv);
    if (dj) cudaFree(dj);
    if (dvalid) cudaFree(dvalid);
    if (rc == 0 && !made && out_choice) out_choice[0] = -1;      // served by a kernel without launch candidates: nothing to tune
    return rc;
}

int skb_plan_get_tuned(const skb_plan *p, const skb_sdf *s, int32_t dtype, int32_t mode, int32_t outputs, int32_t *out_choice) {
    if (!p || !out_choice) return set_error("plan_get_tuned: bad arguments");
    if ((outputs & 7) == 0) return set_error("plan_get_tuned: no output selected");
    if (!s && !p->pairs_set) return set_error("plan_get_tuned: pass an SDF or set the plan's pairs");
    // the launch path up to the kernel selection runs with placeholder (16-byte aligned, never dereferenced) buffers
    void *fake = reinterpret_cast<void *>(uintptr_t(4096));
    s2_tune_begin(1);
    const int rc = s ? skb_collfree(p, s, dtype, fake, 1, 0.0, mode, (outputs & 1) ? fake : nullptr, (outputs & 2) ? fake : nullptr,
                                    (outputs & 4) ? reinterpret_cast<uint8_t *>(fake) : nullptr, nullptr)
                     : skb_pairwise(p, dtype, fake, 1, mode, (outputs & 1) ? fake : nullptr, (outputs & 2) ? fake : nullptr,
                                    (outputs & 4) ? reinterpret_cast<uint8_t *>(fake) : nullptr, nullptr);
    const int found = s2_tune_end(out_choice);
    if (rc != 0) return rc;
    if (!found) out_choice[0] = -1;
    return 0;
}

int skb_com_host(const skb_plan *p_, const skb_sdf *s, int32_t dtype, const void *q_host, int64_t N, int32_t what,
                 void *out_vals_host, void *out_jac_host) {
    skb_plan *p = const_cast<skb_plan *>(p_);
    if (!p || N < 0 || (N > 0 && (!q_host || !out_vals_host))) return set_error("com_host: bad arguments");
    const size_t M = what == SKB_COM_POINT ? 3 : 1;
    return host_pipeline(p, dtype, q_host, N, M, M * p->D, out_vals_host, out_jac_host, nullptr,
                         [&](void *dq, int64_t n, void *dv, void *dj, uint8_t *, cudaStream_t st) {
                             return skb_com(p, s, dtype, dq, n, what, dv, dj, st);
                         });
}

int skb_fk_host(const skb_plan *p_, int32_t dtype, const void *q_host, int64_t N, int32_t with_jacobian, void *out_pts_host,
                void *out_jac_host) {
    skb_plan *p = const_cast<skb_plan *>(p_);
    if (!p || N < 0 || (N > 0 && (!q_host || !out_pts_host))) return set_error("fk_host: bad arguments");
    if (N > 0 && with_jacobian && !out_jac_host) return set_error("fk_host: with_jacobian set but out_jac is NULL");
    const size_t vpc = (size_t)p->F * p->T, jpc = vpc * p->D;
    return host_pipeline(p, dtype, q_host, N, vpc, jpc, out_pts_host, with_jacobian ? out_jac_host : nullptr, nullptr,
                         [&](void *dq, int64_t n, void *dv, void *dj, uint8_t *, cudaStream_t st) {
                             return skb_fk(p, dtype, dq, n, dj != nullptr, dv, dj, st);
                         });
}

int skb_collfree_host(const skb_plan *p_, const skb_sdf *s, int32_t dtype, const void *q_host, int64_t N, double margin,
                      int32_t mode, void *out_vals_host, void *out_jac_host, uint8_t *out_valid_host) {
    skb_plan *p = const_cast<skb_plan *>(p_);
    if (!p || !s || N < 0 || (N > 0 && !q_host)) return set_error("collfree_host: bad arguments");
    const size_t M = mode == SKB_MODE_CLOSEST ? 1 : (size_t)p->F;
    return host_pipeline(p, dtype, q_host, N, M, M * p->D, out_vals_host, out_jac_host, out_valid_host,
                         [&](void *dq, int64_t n, void *dv, void *dj, uint8_t *dvalid, cudaStream_t st) {
                             return skb_collfree(p, s, dtype, dq, n, margin, mode, dv, dj, dvalid, st);
                         });
}

int skb_pairwise_host(const skb_plan *p_, int32_t dtype, const void *q_host, int64_t N, int32_t mode, void *out_vals_host,
                      void *out_jac_host, uint8_t *out_valid_host) {
    skb_plan *p = const_cast<skb_plan *>(p_);
    if (!p || N < 0 || (N > 0 && !q_host)) return set_error("pairwise_host: bad arguments");
    if (!p->pairs_set) return set_error("pairwise: skb_plan_set_pairs was not called");
    const size_t M = mode == SKB_MODE_CLOSEST ? 1 : p->pair_thr.size();
    return host_pipeline(p, dtype, q_host, N, M, M * p->D, out_vals_host, out_jac_host, out_valid_host,
                         [&](void *dq, int64_t n, void *dv, void *dj, uint8_t *dvalid, cudaStream_t st) {
                             return skb_pairwise(p, dtype, dq, n, mode, dv, dj, dvalid, st);
                         });
}

}  // extern "C"

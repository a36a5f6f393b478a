// C-ABI glue: error reporting, the final reductions, the whole-path orchestrators (device and host
// buffers) and the host-side segment builder.  See include/svgpfa_b200.h.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

int svgpfa_launch_spike_gather(const svgpfa_dims* dims, const svgpfa_buffers* buf, bool reuse, cudaStream_t stream);
bool svgpfa_try_chol_indpoints_fused(const svgpfa_dims* dims, const svgpfa_buffers* buf, cudaStream_t st);

namespace {

thread_local char g_err[256] = "";
thread_local cudaEvent_t* g_stage_events = nullptr;

inline void stage_mark(int idx, cudaStream_t st) {
    if (g_stage_events) cudaEventRecord(g_stage_events[idx], st);
}

// ------------------------------------------------------------------------------------------
// reductions into `shared`
//   shared[2] = KL, shared[3] = term1, shared[4] += alpha . abar_spk + cnt . d, dtheta = sum_r dth_part,
//   dd += spike counts
// ------------------------------------------------------------------------------------------
constexpr int FIN_THREADS = 256;

// Per-block partials go to fin_part[3][SVGPFA_FIN_SLOTS] and are summed in slot order by finalize_combine_kernel: no
// floating-point atomics here, so the three scalars are reproducible run to run given the stage buffers.
__global__ void __launch_bounds__(FIN_THREADS) finalize_reduce_kernel(svgpfa_dims dm, svgpfa_buffers bf, uint32_t flags,
                                                                      int use_abar) {
    __shared__ double red[32];
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gstr = (size_t)gridDim.x * blockDim.x;
    double kl = 0.0, t1 = 0.0, t2 = 0.0;
    for (size_t i = gtid; i < (size_t)dm.R * dm.K; i += gstr) kl += bf.kl_rk[i];
    for (size_t i = gtid; i < SVGPFA_TERM1_SLOTS; i += gstr) t1 += bf.term1_part[i];
    if (use_abar)
        for (size_t i = gtid; i < (size_t)dm.R * dm.KM; i += gstr) t2 = fma(bf.alpha[i], bf.abar_spk[i], t2);
    double* gd = bf.shared + SVGPFA_SHARED_HDR + (size_t)dm.N * dm.K;
    for (size_t n = gtid; n < (size_t)dm.N; n += gstr) {
        t2 = fma(bf.spike_cnt[n], bf.d[n], t2);
        if (flags & SVGPFA_GRAD_EMBEDDING) gd[n] += bf.spike_cnt[n];
    }
    const double skl = block_sum(kl, red);
    const double st1 = block_sum(t1, red);
    const double st2 = block_sum(t2, red);
    if (threadIdx.x == 0) {
        bf.fin_part[blockIdx.x] = skl;
        bf.fin_part[SVGPFA_FIN_SLOTS + blockIdx.x] = st1;
        bf.fin_part[2 * SVGPFA_FIN_SLOTS + blockIdx.x] = st2;
    }
    if (flags & SVGPFA_GRAD_KERNEL) {
        // dtheta[i] = sum_r dth_part[r][i]: one warp per parameter, lanes stride over trials
        double* gth = gd + dm.N;
        const int wid = (int)(gtid >> 5), lane = threadIdx.x & 31, nw = (int)(gstr >> 5);
        for (int i = wid; i < dm.TH; i += nw) {
            double s = 0.0;
            for (int r = lane; r < dm.R; r += 32) s += bf.dth_part[(size_t)r * dm.TH + i];
            s = warp_sum(s);
            if (lane == 0) gth[i] = s;
        }
    }
}

// one warp: lane l sums slots l, l + 32, ... in order, then a fixed shuffle tree.  shared[4] may already hold the
// C part of term2 (cached-statistics path: written by the gather kernel), hence "+=".
__global__ void finalize_combine_kernel(svgpfa_buffers bf, int nslots, int with_kl) {
    const int lane = threadIdx.x;
    double v[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double s = 0.0;
        for (int i = lane; i < nslots; i += 32) s += bf.fin_part[c * SVGPFA_FIN_SLOTS + i];
        v[c] = warp_sum(s);
    }
    if (lane == 0) {
        double* s = bf.shared;
        s[2] = with_kl ? v[0] : 0.0;
        s[3] = v[1];
        s[4] += v[2];
        const double ell = -s[3] + s[4];
        s[1] = ell;
        s[0] = with_kl ? ell - s[2] : ell;
        s[5] = (double)bf.info[0];
        s[6] = (double)bf.info[1];
        s[7] = (double)bf.info[2];
    }
}

// every accumulator of one evaluation in ONE launch (was six memset nodes)
struct ZeroList {
    double* p[6];
    size_t n[6];
    int cnt;
};

__global__ void __launch_bounds__(256) zero_kernel(ZeroList zl) {
    const size_t gtid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, gstr = (size_t)gridDim.x * blockDim.x;
    for (int a = 0; a < zl.cnt; ++a) {
        double* p = zl.p[a];
        const size_t n = zl.n[a];
        for (size_t i = gtid; i < n; i += gstr) p[i] = 0.0;
    }
}

int check_dims(const svgpfa_dims* d, const svgpfa_buffers* b, const char* where) {
    if (!d || !b) return svgpfa_set_error(SVGPFA_E_ARG, where, cudaSuccess);
    if (d->R < 0 || d->N < 0 || d->K < 1 || d->Q < 0 || d->Mmax < 1 || d->Mmax > SVGPFA_MAX_M)
        return svgpfa_set_error(SVGPFA_E_ARG, where, cudaSuccess);
    return SVGPFA_OK;
}

size_t shared_len(const svgpfa_dims* d) { return SVGPFA_SHARED_HDR + (size_t)d->N * d->K + d->N + d->TH; }

}  // namespace

int svgpfa_sm_count() {
    static std::atomic<int> cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    int n = cache[dev & 63].load(std::memory_order_relaxed);
    if (n <= 0) {
        n = 148;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        cache[dev & 63].store(n, std::memory_order_relaxed);
    }
    return n;
}

int svgpfa_set_error(int code, const char* where, cudaError_t ce) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, ce == cudaSuccess ? (code == SVGPFA_E_ARG ? "bad argument" : "unsupported") : cudaGetErrorString(ce));
    return code;
}

extern "C" int svgpfa_abi_version(void) { return SVGPFA_ABI_VERSION; }

extern "C" int svgpfa_set_stage_events(void** events) {
    g_stage_events = reinterpret_cast<cudaEvent_t*>(events);
    return SVGPFA_OK;
}

extern "C" const char* svgpfa_last_error(void) { return g_err; }

static int finalize_blocks(const svgpfa_dims* dims) {
    size_t work = (size_t)dims->R * dims->KM;
    int blocks = (int)((work + FIN_THREADS * 8 - 1) / (FIN_THREADS * 8));
    if (blocks < 1) blocks = 1;
    if (blocks > 592) blocks = 592;
    return blocks;
}

extern "C" int svgpfa_finalize(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, void* stream) {
    int rc = check_dims(dims, buf, "finalize");
    if (rc) return rc;
    if (!buf->fin_part) return svgpfa_set_error(SVGPFA_E_ARG, "finalize: fin_part", cudaSuccess);
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = finalize_blocks(dims);
    finalize_reduce_kernel<<<blocks, FIN_THREADS, 0, st>>>(*dims, *buf, flags, 1);
    SVGPFA_CHECK_LAUNCH("finalize_reduce");
    finalize_combine_kernel<<<1, 32, 0, st>>>(*buf, blocks, 1);
    SVGPFA_CHECK_LAUNCH("finalize_combine");
    return SVGPFA_OK;
}

namespace {

// zero every accumulator of one evaluation (whole shard)
void zero_accumulators(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, bool reuse_spike, cudaStream_t st) {
    const uint32_t kz = SVGPFA_GRAD_KERNEL | SVGPFA_GRAD_INDLOCS;
    ZeroList zl;
    zl.cnt = 0;
    auto add = [&](void* p, size_t n) { if (p && n) { zl.p[zl.cnt] = (double*)p; zl.n[zl.cnt] = n; ++zl.cnt; } };
    add(buf->shared, shared_len(dims));
    add(buf->term1_part, SVGPFA_TERM1_SLOTS);
    add(buf->info, 2);                                  // 4 x int32
    size_t big = 0;
    if (!reuse_spike) { add(buf->abar_spk, (size_t)dims->R * dims->KM); big += (size_t)dims->R * dims->KM; }
    if (flags & kz) {
        add(buf->dz_acc, (size_t)dims->R * dims->KM);
        add(buf->dth_part, (size_t)dims->R * dims->TH);
        big += (size_t)dims->R * (dims->KM + dims->TH);
    }
    size_t blocks = (big + shared_len(dims) + 256 * 8 - 1) / (256 * 8);
    if (blocks < 1) blocks = 1;
    if (blocks > 1184) blocks = 1184;
    zero_kernel<<<(unsigned)blocks, 256, 0, st>>>(zl);
}

// the seven per-trial stages on the trial range of `dims` (r0, rn); stage events only when `mark`
int run_trial_stages(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, bool reuse_spike,
                     cudaStream_t st, bool mark) {
    const uint32_t lat = SVGPFA_GRAD_POSTERIOR | SVGPFA_GRAD_KERNEL | SVGPFA_GRAD_INDLOCS;
    void* stream = (void*)st;
    int rc;
    // (Running the quadrature chain on a high-priority side stream concurrently with the spike kernel was measured:
    //  32.41 vs 32.53 ms on the 2000-trial shard -- both kernels are limited by the same FP64 issue port and the spike
    //  kernel alone already holds every register of the SM, so the stages simply run one after the other.)
    if (mark) stage_mark(0, st);
    bool fused = false;
    if (!(flags & SVGPFA_REUSE_KZZ)) {
        // M <= 32: Cholesky, inverse and the X / c / alpha / KL stage in one launch (the stage event of kzz_chol then
        // covers both; indpoints_fwd reads 0)
        fused = svgpfa_try_chol_indpoints_fused(dims, buf, st);
        if (fused) { SVGPFA_CHECK_LAUNCH("kzz_chol + indpoints_fwd"); }
        else { rc = svgpfa_kzz_chol_fwd(dims, buf, stream); if (rc) return rc; }
    }
    if (mark) stage_mark(1 + SVGPFA_STAGE_KZZ_CHOL, st);
    if (!fused) { rc = svgpfa_indpoints_fwd(dims, buf, stream); if (rc) return rc; }
    if (mark) stage_mark(1 + SVGPFA_STAGE_INDPOINTS_FWD, st);
    rc = (flags & SVGPFA_REUSE_VQ) ? svgpfa_quad_latent_fwd_cached(dims, buf, stream) : svgpfa_quad_latent_fwd(dims, buf, stream);
    if (rc) return rc;
    if (mark) stage_mark(1 + SVGPFA_STAGE_QUAD_LATENT_FWD, st);
    rc = svgpfa_quad_embed_fwd_bwd(dims, buf, flags, stream); if (rc) return rc;
    if (mark) stage_mark(1 + SVGPFA_STAGE_QUAD_EMBED, st);
    if (flags & lat) { rc = svgpfa_quad_latent_bwd(dims, buf, flags, stream); if (rc) return rc; }
    if (mark) stage_mark(1 + SVGPFA_STAGE_QUAD_LATENT_BWD, st);
    if (!reuse_spike) {
        rc = dims->spike_method == SVGPFA_SPIKE_PANEL ? svgpfa_spike_panel_fwd_bwd(dims, buf, flags, stream)
                                                      : svgpfa_spike_fwd_bwd(dims, buf, flags, stream);
        if (rc) return rc;
    }
    if (mark) stage_mark(1 + SVGPFA_STAGE_SPIKE, st);
    if (flags & lat) { rc = svgpfa_indpoints_bwd(dims, buf, flags, stream); if (rc) return rc; }
    if (mark) stage_mark(1 + SVGPFA_STAGE_INDPOINTS_BWD, st);
    return SVGPFA_OK;
}

}  // namespace

extern "C" int svgpfa_elbo_grad(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, void* stream) {
    int rc = check_dims(dims, buf, "elbo_grad");
    if (rc) return rc;
    if (dims->r0 != 0 || (dims->rn != 0 && dims->rn != dims->R))
        return svgpfa_set_error(SVGPFA_E_ARG, "elbo_grad: a trial range is only valid on the per-stage entry points", cudaSuccess);
    cudaStream_t st = (cudaStream_t)stream;
    const uint32_t kz = SVGPFA_GRAD_KERNEL | SVGPFA_GRAD_INDLOCS;
    const bool reuse_spike = (flags & SVGPFA_REUSE_SPIKE) && !(flags & (kz | SVGPFA_GRAD_EMBEDDING));
    zero_accumulators(dims, buf, flags, reuse_spike, st);
    // (Trial blocks alternating between two streams were measured here too: 263.2 vs 264.3 ms at config #5 -- the
    //  kernels of one evaluation already fill the machine; only the pipelined host-buffer entry gains from it.)
    rc = run_trial_stages(dims, buf, flags, reuse_spike, st, true);
    if (rc) return rc;
    rc = svgpfa_finalize(dims, buf, flags, stream);
    stage_mark(1 + SVGPFA_STAGE_FINALIZE, st);
    return rc;
}

extern "C" int svgpfa_cached_ell_fwd_bwd(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, void* stream) {
    int rc = check_dims(dims, buf, "cached_ell_fwd_bwd");
    if (rc) return rc;
    if (!buf->fin_part) return svgpfa_set_error(SVGPFA_E_ARG, "cached_ell_fwd_bwd: fin_part", cudaSuccess);
    cudaStream_t st = (cudaStream_t)stream;
    ZeroList zl;
    zl.cnt = 2;
    zl.p[0] = buf->shared; zl.n[0] = shared_len(dims);
    zl.p[1] = buf->term1_part; zl.n[1] = SVGPFA_TERM1_SLOTS;
    zero_kernel<<<8, 256, 0, st>>>(zl);
    rc = svgpfa_quad_embed_fwd_bwd(dims, buf, SVGPFA_GRAD_EMBEDDING, stream); if (rc) return rc;
    rc = svgpfa_launch_spike_gather(dims, buf, (flags & SVGPFA_REUSE_SPIKE) != 0, st); if (rc) return rc;
    // term1 and the d part of term2; no KL, no alpha.abar (the gather produced the C part in shared[4])
    svgpfa_dims d0 = *dims;
    d0.R = 0;                                   // empties the kl_rk loop; the alpha.abar loop is off (use_abar = 0)
    finalize_reduce_kernel<<<8, FIN_THREADS, 0, st>>>(d0, *buf, SVGPFA_GRAD_EMBEDDING, 0);
    SVGPFA_CHECK_LAUNCH("cached finalize_reduce");
    finalize_combine_kernel<<<1, 32, 0, st>>>(*buf, 8, 0);
    SVGPFA_CHECK_LAUNCH("cached finalize_combine");
    return SVGPFA_OK;
}

extern "C" int svgpfa_build_segments_host(int32_t R, int32_t N, const int64_t* counts_host, int64_t* seg_off_host,
                                          int64_t* neuron_index_host) {
    if (R < 0 || N < 0 || !counts_host || !seg_off_host) return svgpfa_set_error(SVGPFA_E_ARG, "build_segments_host", cudaSuccess);
    int64_t acc = 0;
    seg_off_host[0] = 0;
    for (int64_t i = 0; i < (int64_t)R * N; ++i) {
        const int64_t c = counts_host[i];
        if (c < 0) return svgpfa_set_error(SVGPFA_E_ARG, "build_segments_host: negative count", cudaSuccess);
        if (neuron_index_host) {
            const int64_t n = i % N;
            for (int64_t s = 0; s < c; ++s) neuron_index_host[acc + s] = n;
        }
        acc += c;
        seg_off_host[i + 1] = acc;
    }
    return SVGPFA_OK;
}

namespace {

// Side streams and events of the pipelined host-buffer entry, one set per (host thread, device), created on first use.
constexpr int HP_MAX_BLOCKS = 16;
struct HostPipe {
    bool init = false;
    cudaStream_t in = nullptr, out = nullptr, comp2 = nullptr;
    cudaEvent_t start = nullptr, drained = nullptr, comp2_done = nullptr, copied[HP_MAX_BLOCKS], done[HP_MAX_BLOCKS];
};
thread_local HostPipe g_pipes[16];

HostPipe* host_pipe() {
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16) return nullptr;
    HostPipe& p = g_pipes[dev];
    if (!p.init) {
        if (cudaStreamCreateWithFlags(&p.in, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaStreamCreateWithFlags(&p.out, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaStreamCreateWithFlags(&p.comp2, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        cudaEventCreateWithFlags(&p.comp2_done, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&p.start, cudaEventDisableTiming);
        cudaEventCreateWithFlags(&p.drained, cudaEventDisableTiming);
        for (int i = 0; i < HP_MAX_BLOCKS; ++i) {
            cudaEventCreateWithFlags(&p.copied[i], cudaEventDisableTiming);
            cudaEventCreateWithFlags(&p.done[i], cudaEventDisableTiming);
        }
        p.init = true;
    }
    return &p;
}

}  // namespace

extern "C" int svgpfa_release_thread_resources(void) {
    int cur = 0;
    cudaGetDevice(&cur);
    for (int dev = 0; dev < 16; ++dev) {
        HostPipe& p = g_pipes[dev];
        if (!p.init) continue;
        cudaSetDevice(dev);
        cudaStreamDestroy(p.in);
        cudaStreamDestroy(p.out);
        cudaStreamDestroy(p.comp2);
        cudaEventDestroy(p.comp2_done);
        cudaEventDestroy(p.start);
        cudaEventDestroy(p.drained);
        for (int i = 0; i < HP_MAX_BLOCKS; ++i) {
            cudaEventDestroy(p.copied[i]);
            cudaEventDestroy(p.done[i]);
        }
        p = HostPipe();
    }
    cudaSetDevice(cur);
    g_stage_events = nullptr;
    return SVGPFA_OK;
}

extern "C" int svgpfa_elbo_grad_host(const svgpfa_dims* dims, const svgpfa_buffers* dev, const svgpfa_host_io* io,
                                     uint32_t flags, void* stream) {
    int rc = check_dims(dims, dev, "elbo_grad_host");
    if (rc) return rc;
    if (!io) return svgpfa_set_error(SVGPFA_E_ARG, "elbo_grad_host", cudaSuccess);
    if (!dims->desc_host) return svgpfa_set_error(SVGPFA_E_ARG, "elbo_grad_host: dims.desc_host", cudaSuccess);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t R = dims->R, D = sizeof(double);
    uint32_t flags_run = flags & ~(uint32_t)(SVGPFA_REUSE_KZZ | SVGPFA_REUSE_SPIKE | SVGPFA_REUSE_VQ);   // every input is new
    if (io->copy_static && dims->spike_method == SVGPFA_SPIKE_PANEL) flags_run |= SVGPFA_REBUILD_PANELS;   // new spikes
#define CPY(dst, src, bytes, kind, s_)                                                                \
    do {                                                                                              \
        if ((bytes) > 0 && (dst) && (src)) {                                                          \
            cudaError_t e = cudaMemcpyAsync((void*)(dst), (src), (bytes), kind, s_);                  \
            if (e != cudaSuccess) return svgpfa_set_error(SVGPFA_E_CUDA, "elbo_grad_host copy", e);   \
        }                                                                                             \
    } while (0)
#define H2D(dst, src, bytes, s_) CPY(dst, src, bytes, cudaMemcpyHostToDevice, s_)
#define D2H(dst, src, bytes, s_) CPY(dst, src, bytes, cudaMemcpyDeviceToHost, s_)
    // blocks of trials: copies of block b+1 (copy-in stream) and of block b-1's gradients (copy-out stream) run
    // under the kernels of block b (caller's stream); 1 block = everything in order on the caller's stream
    int nb = io->n_blocks;
    if (nb <= 0) nb = (int)(R / 312);           // measured (tools/bench_host.py): 2500 trials 50.4 / 41.8 / 41.0 ms with
                                                // 1 / 4 / 8 blocks; 20000 trials 324.6 / 314.1 / 308.7 ms with 4 / 8 / 16
                                                // (32 blocks: 336 ms -- launch and copy calls start to dominate)
    if (nb > HP_MAX_BLOCKS) nb = HP_MAX_BLOCKS;
    if (nb > (int)R) nb = (int)R;
    if (nb < 1) nb = 1;
    if (io->copy_static && (!io->seg_off_host || !io->spike_t_host))
        return svgpfa_set_error(SVGPFA_E_ARG, "elbo_grad_host: copy_static needs seg_off_host and spike_t_host", cudaSuccess);
    HostPipe* hp = nb > 1 ? host_pipe() : nullptr;
    if (!hp) nb = 1;
    cudaStream_t s_in = hp ? hp->in : st, s_out = hp ? hp->out : st;
    H2D(dev->theta, io->theta_host, (size_t)dims->TH * D, st);
    H2D(dev->C, io->C_host, (size_t)dims->N * dims->K * D, st);
    H2D(dev->d, io->d_host, (size_t)dims->N * D, st);
    if (io->copy_static) H2D(dev->spike_cnt, io->spike_cnt_host, (size_t)dims->N * D, st);
    zero_accumulators(dims, dev, flags_run, false, st);
    if (hp) {
        cudaEventRecord(hp->start, st);            // earlier work on the caller's stream may still read the buffers
        cudaStreamWaitEvent(s_in, hp->start, 0);
    }
    // Odd blocks run their kernels on a second compute stream, so that the tail wave of one block's kernels overlaps
    // the next block's (20000 trials: 274.0 -> 269.6 ms, 2500 trials: 36.3 -> 35.4 ms).  Both streams accumulate into `shared` with atomics; per-trial outputs never overlap.
    const bool two = hp != nullptr;
    if (two) cudaStreamWaitEvent(hp->comp2, hp->start, 0);
    // Block boundaries.  The copy-in of the first block and the copy-out of the last one cannot hide under kernels, so
    // the blocks at both ends are small and grow towards the middle: weights 1, 2, ..., 8, 8, ..., 2, 1.
    size_t bw_prefix[HP_MAX_BLOCKS + 1];
    bw_prefix[0] = 0;
    for (int b = 0; b < nb; ++b) {
        int w = b + 1 < nb - b ? b + 1 : nb - b;
        if (w > 8) w = 8;
        bw_prefix[b + 1] = bw_prefix[b] + (size_t)w;
    }
    auto r_of = [&](int b) { return (size_t)(R * bw_prefix[b] / bw_prefix[nb]); };
    // K-major arrays: the block's trials are K separate runs, one per latent; with a uniform M they form a 2-D copy
    bool uniform = true;
    for (int k = 1; k < dims->K; ++k) uniform = uniform && dims->desc_host[k].M == dims->desc_host[0].M;
    auto kmajor = [&](void* d, const void* h, bool per_p, size_t r0, size_t n, cudaMemcpyKind kind, cudaStream_t s_) -> cudaError_t {
        if (!d || !h || n == 0) return cudaSuccess;
        char* dc = (char*)d;
        const char* hc = (const char*)h;
        if (uniform) {
            const size_t w = per_p ? dims->desc_host[0].P : dims->desc_host[0].M;
            return cudaMemcpy2DAsync(dc + r0 * w * D, R * w * D, hc + r0 * w * D, R * w * D, n * w * D, dims->K, kind, s_);
        }
        for (int k = 0; k < dims->K; ++k) {
            const svgpfa_latent_desc& ds = dims->desc_host[k];
            const size_t w = per_p ? ds.P : ds.M, o = (R * (per_p ? ds.poff : ds.moff) + r0 * w) * D;
            cudaError_t e = cudaMemcpyAsync(dc + o, hc + o, n * w * D, kind, s_);
            if (e != cudaSuccess) return e;
        }
        return cudaSuccess;
    };
#define KM(call)                                                                                      \
    do {                                                                                              \
        cudaError_t e = (call);                                                                       \
        if (e != cudaSuccess) return svgpfa_set_error(SVGPFA_E_CUDA, "elbo_grad_host block copy", e); \
    } while (0)
    auto copy_in = [&](int b) -> int {
        const size_t r0 = r_of(b), r1 = r_of(b + 1), n = r1 - r0;
        KM(kmajor((void*)dev->Z, io->Z_host, false, r0, n, cudaMemcpyHostToDevice, s_in));
        KM(kmajor((void*)dev->m, io->m_host, false, r0, n, cudaMemcpyHostToDevice, s_in));
        KM(kmajor((void*)dev->cholvec, io->cholvec_host, true, r0, n, cudaMemcpyHostToDevice, s_in));
        if (io->copy_static) {
            H2D(dev->tq + r0 * dims->Q, io->tq_host + r0 * dims->Q, n * dims->Q * D, s_in);
            H2D(dev->wq + r0 * dims->Q, io->wq_host + r0 * dims->Q, n * dims->Q * D, s_in);
            const size_t g0 = r0 * dims->N, g1 = r1 * dims->N;
            H2D(dev->seg_off + g0, io->seg_off_host + g0, (g1 - g0 + 1) * sizeof(int64_t), s_in);
            const int64_t s0 = io->seg_off_host[g0], s1 = io->seg_off_host[g1];
            H2D(dev->spike_t + s0, io->spike_t_host + s0, (size_t)(s1 - s0) * D, s_in);
        }
        if (hp) cudaEventRecord(hp->copied[b], s_in);
        return SVGPFA_OK;
    };
    // the copies of block b + 1 are enqueued before the kernels of block b, so the host never runs far ahead of
    // the device with copy calls only
    rc = copy_in(0);
    if (rc) return rc;
    for (int b = 0; b < nb; ++b) {
        const size_t r0 = r_of(b), r1 = r_of(b + 1), n = r1 - r0;
        if (b + 1 < nb) { rc = copy_in(b + 1); if (rc) return rc; }
        svgpfa_dims db = *dims;
        db.r0 = (int32_t)r0;
        db.rn = nb > 1 ? (int32_t)n : 0;
        cudaStream_t sc = (two && (b & 1)) ? hp->comp2 : st;
        if (hp) cudaStreamWaitEvent(sc, hp->copied[b], 0);
        if (n > 0) {
            rc = run_trial_stages(&db, dev, flags_run, false, sc, nb == 1);
            if (rc) return rc;
        }
        if (hp) {
            cudaEventRecord(hp->done[b], sc);
            cudaStreamWaitEvent(s_out, hp->done[b], 0);
        }
        if (nb > 1) {                              // this block's per-trial gradients
            if (flags & SVGPFA_GRAD_INDLOCS) KM(kmajor(io->gZ_host, dev->gZ, false, r0, n, cudaMemcpyDeviceToHost, s_out));
            if (flags & SVGPFA_GRAD_POSTERIOR) {
                KM(kmajor(io->gm_host, dev->gm, false, r0, n, cudaMemcpyDeviceToHost, s_out));
                KM(kmajor(io->gcholvec_host, dev->gcholvec, true, r0, n, cudaMemcpyDeviceToHost, s_out));
            }
        }
    }
#undef KM
    if (two) {
        cudaEventRecord(hp->comp2_done, hp->comp2);
        cudaStreamWaitEvent(st, hp->comp2_done, 0);
    }
    rc = svgpfa_finalize(dims, dev, flags_run, stream);
    if (nb == 1) stage_mark(1 + SVGPFA_STAGE_FINALIZE, st);
    if (rc) return rc;
    D2H(io->shared_host, dev->shared, shared_len(dims) * D, st);
    D2H(io->info_host, dev->info, 4 * sizeof(int32_t), st);
    if (nb == 1) {
        if (flags & SVGPFA_GRAD_INDLOCS) D2H(io->gZ_host, dev->gZ, R * dims->KM * D, st);
        if (flags & SVGPFA_GRAD_POSTERIOR) {
            D2H(io->gm_host, dev->gm, R * dims->KM * D, st);
            D2H(io->gcholvec_host, dev->gcholvec, R * dims->PP * D, st);
        }
    } else {
        cudaEventRecord(hp->drained, s_out);       // the caller's stream drains only after the last gradient copy
        cudaStreamWaitEvent(st, hp->drained, 0);
    }
#undef CPY
#undef H2D
#undef D2H
    return SVGPFA_OK;
}

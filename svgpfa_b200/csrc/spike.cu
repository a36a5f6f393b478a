// Spike-time (AssocTimes) term of the expected log-likelihood over ragged CSR spike trains:
//   spike_fwd_bwd_kernel     sum_s [ sum_k C[n_s,k] kappa_k(t_s, Z_kr) . alpha_kr ] and every adjoint in one pass
//   spike_means_kernel       mu_s[s][k] = kappa_k(t_s, Z_kr) . alpha_kr           (cached-statistics path)
//   spike_gather_kernel      sum_s sum_k mu_s[s][k] C[n_s,k], dC[n][k] += segment sums  (HBM-bound gather)
// Replaces stats/kernelsMatricesStore.py:208-221, stats/svPosteriorOnLatents.py:265-300 (mean only: the
// exponential link never uses the spike-time variance, expectedLogLikelihood.py:210-213),
// stats/svEmbedding.py:137-144.  Ktz[k][r] (S_r x M) is never materialised.
//
// Mapping of spike_fwd_bwd_kernel: a warp owns 32 consecutive (latent, inducing point) pairs of one trial
// -- z_j, alpha_j and the kernel constants live in registers -- and walks the spikes of a range of neurons.
// Spikes are stored neuron-major inside a trial, so the embedding weight C[n,k] is constant over a segment:
//   pn_j(n)   = sum_{s in (r,n)} kappa(t_s - z_j)
//   abar_j    = sum_n C[n,k] pn_j(n)          (no cross-lane traffic)
//   dC[n,k]  += sum_j alpha_j pn_j(n)         (one segmented warp reduction per NON-EMPTY segment)
// and the value of the term is alpha . abar (taken in svgpfa_finalize).
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int SP_WPB = 4;     // warps per CTA

// segmented (by latent) sum over the lanes of a warp; valid in the first lane of every segment
__device__ __forceinline__ double seg_sum(double v, unsigned same_mask) {
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        const double o = __shfl_down_sync(0xffffffffu, v, 1 << b);
        if (same_mask & (1u << b)) v += o;
    }
    return v;
}

// One kernel evaluation and its accumulations for one (spike, inducing point).
//   pn += kappa;  expquad : p1 += kappa d,              p2 += kappa d^2
//                 periodic: p1 += kappa sin(2 pi d/p),  p2 += kappa sin^2(pi d/p),  p3 += kappa sin(2 pi d/p) d
template <bool KGRAD, bool PERIODIC>
__device__ __forceinline__ void spike_eval(double t, double z, double nh, double invp,
                                           const double* __restrict__ etab, double& pn, double& p1, double& p2,
                                           double& p3) {
    const double dl = t - z;
    if (!PERIODIC) {
        const double q = dl * dl;
        const double kv = svgpfa_exp_neg(nh * q, etab);
        pn += kv;
        if (KGRAD) {
            p1 = fma(kv, dl, p1);
            p2 = fma(kv, q, p2);
        }
    } else {
        double sn, cs;
        sincospi(dl * invp, &sn, &cs);
        const double q = sn * sn;
        const double kv = svgpfa_exp_neg(nh * q, etab);
        pn += kv;
        if (KGRAD) {
            const double w = kv * (2.0 * sn * cs);
            p1 += w;
            p2 = fma(kv, q, p2);
            p3 = fma(w, dl, p3);
        }
    }
}

// All spikes of one (trial, neuron) segment for one lane.  UNROLL spike times are loaded one iteration ahead
// of their use (software pipelining: the loads are warp-uniform L1 hits, but their latency would otherwise
// sit in front of every dependent FP64 chain).
template <bool KGRAD, bool PERIODIC, int UNROLL>
__device__ __forceinline__ void spike_segment(const double* __restrict__ sp, int cnt, double z, double nh, double invp,
                                              const double* __restrict__ etab, double& pn, double& p1, double& p2,
                                              double& p3) {
    int i = 0;
    if (cnt >= UNROLL) {
        double tn[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) tn[u] = sp[u];
        for (; i + UNROLL <= cnt; i += UNROLL) {
            double tc[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) tc[u] = tn[u];
            if (i + 2 * UNROLL <= cnt) {
#pragma unroll
                for (int u = 0; u < UNROLL; ++u) tn[u] = sp[i + UNROLL + u];
            }
            // pull the line 32 spikes ahead into L1 (segments of a trial are contiguous, so this also warms the
            // next segments); without it every fourth iteration waits on an L2 round trip
            asm volatile("prefetch.global.L1 [%0];" ::"l"(sp + i + 32));
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) spike_eval<KGRAD, PERIODIC>(tc[u], z, nh, invp, etab, pn, p1, p2, p3);
        }
    }
    for (; i < cnt; ++i) spike_eval<KGRAD, PERIODIC>(sp[i], z, nh, invp, etab, pn, p1, p2, p3);
}

// UNROLL = spikes per software-pipelined iteration, MINB = resident CTAs per SM asked of ptxas.
template <bool KGRAD, int UNROLL, int MINB>
__global__ void __launch_bounds__(32 * SP_WPB, MINB) spike_fwd_bwd_kernel(svgpfa_dims dm, svgpfa_buffers bf,
                                                                          uint32_t flags, int n_chunks, int chunk) {
    __shared__ double etab[SVGPFA_EXP_TAB_SIZE];
    svgpfa_load_exp_tab(etab);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r = blockIdx.x / n_chunks, nc = blockIdx.x - r * n_chunks;
    const int grp = blockIdx.y * SP_WPB + warp;
    const int li = grp * 32 + lane;
    if (grp * 32 >= dm.KM) return;                       // whole warp out of range
    const bool active = li < dm.KM;
    // (latent, inducing point) of this lane
    int k = 0;
    const int l = active ? li : dm.KM - 1;
    while (k + 1 < dm.K && bf.desc[k + 1].moff <= l) ++k;
    const double z = bf.Z[(size_t)dm.R * bf.desc[k].moff + (size_t)r * bf.desc[k].M + (l - bf.desc[k].moff)];
    double nh, invp, s2;
    bool periodic;
    {
        const KConst kc = make_kconst(bf.desc[k], bf.theta, bf.kscale, k);
        nh = kc.nh; invp = kc.invp; s2 = kc.s2; periodic = kc.type == SVGPFA_KERNEL_PERIODIC;
    }
    // segment structure of the warp (lanes of one latent are contiguous)
    unsigned same = 0;
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        const int ok = __shfl_down_sync(0xffffffffu, k, 1 << b);
        const int oact = __shfl_down_sync(0xffffffffu, (int)active, 1 << b);
        if (lane + (1 << b) < 32 && ok == k && oact && active) same |= 1u << b;
    }
    const int kprev = __shfl_up_sync(0xffffffffu, k, 1);
    const bool head = active && (lane == 0 || kprev != k);
    const bool need_emb = flags & SVGPFA_GRAD_EMBEDDING;
    const double a = active ? s2 * bf.alpha[(size_t)r * dm.KM + li] : 0.0;     // scale^2 alpha_j

    const int nb = nc * chunk, ne = min(dm.N, nb + chunk);
    const int64_t* __restrict__ seg = bf.seg_off + (size_t)r * dm.N;
    const double* __restrict__ st = bf.spike_t;
    const double* __restrict__ Ck = bf.C + k;
    double* gCk = bf.shared + SVGPFA_SHARED_HDR + k;
    double abar = 0.0, dz = 0.0, d0 = 0.0, d1 = 0.0;
    int64_t s1 = seg[nb];
    for (int n = nb; n < ne; ++n) {
        const int64_t s0 = s1;
        s1 = seg[n + 1];
        const int cnt = (int)(s1 - s0);
        if (cnt == 0) continue;
        const double c = Ck[(size_t)n * dm.K];
        double pn = 0.0, p1 = 0.0, p2 = 0.0, p3 = 0.0;
        if (!periodic) spike_segment<KGRAD, false, UNROLL>(st + s0, cnt, z, nh, invp, etab, pn, p1, p2, p3);
        else spike_segment<KGRAD, true, UNROLL>(st + s0, cnt, z, nh, invp, etab, pn, p1, p2, p3);
        abar = fma(c, pn, abar);
        if (KGRAD) {
            dz = fma(c, p1, dz);
            d0 = fma(c, p2, d0);
            d1 = fma(c, p3, d1);
        }
        if (need_emb) {
            const double v = seg_sum(pn * a, same);
            if (head) atomicAdd(gCk + (size_t)n * dm.K, v);
        }
    }
    const svgpfa_latent_desc ds = bf.desc[k];
    const KConst kc = make_kconst(ds, bf.theta, bf.kscale, k);
    if (active) {
        atomicAdd(bf.abar_spk + (size_t)r * dm.KM + li, kc.s2 * abar);
        // kbar_sj = C alpha_j ; d delta/dz = -1 ; dkappa/ddelta = kappa * (delta | sin 2 pi d/p) * dd
        if (KGRAD && (flags & SVGPFA_GRAD_INDLOCS))
            atomicAdd(bf.dz_acc + (size_t)r * dm.KM + li, -a * kc.dd * dz);
    }
    if (KGRAD && (flags & SVGPFA_GRAD_KERNEL)) {
        // dkappa/dtheta0 = kappa (d^2 | sin^2) dl (p2);  periodic: dkappa/dtheta1 = kappa sin(2 pi d/p) d dp (p3)
        const double t0 = seg_sum(active ? a * kc.dl * d0 : 0.0, same);
        const double t1 = seg_sum(active ? a * kc.dp * d1 : 0.0, same);
        if (head) {
            double* dth = bf.dth_part + (size_t)r * dm.TH + ds.thoff;
            atomicAdd(dth, t0);
            if (ds.nth > 1) atomicAdd(dth + 1, t1);
        }
    }
}

// ------------------------------------------------------------------------------------------
constexpr int SM_THREADS = 256;

__global__ void __launch_bounds__(SM_THREADS) spike_means_kernel(svgpfa_dims dm, svgpfa_buffers bf, int n_split) {
    extern __shared__ double sm[];
    double* zs = sm;                 // KM
    double* as = zs + dm.KM;         // KM
    const int r = blockIdx.x / n_split, part = blockIdx.x - r * n_split;
    for (int k = 0; k < dm.K; ++k) {
        const svgpfa_latent_desc ds = bf.desc[k];
        for (int j = threadIdx.x; j < ds.M; j += blockDim.x) {
            zs[ds.moff + j] = bf.Z[(size_t)dm.R * ds.moff + (size_t)r * ds.M + j];
            as[ds.moff + j] = bf.alpha[(size_t)r * dm.KM + ds.moff + j];
        }
    }
    __syncthreads();
    const int64_t s0 = bf.seg_off[(size_t)r * dm.N], s1 = bf.seg_off[(size_t)(r + 1) * dm.N];
    for (int64_t s = s0 + (int64_t)part * blockDim.x + threadIdx.x; s < s1; s += (int64_t)n_split * blockDim.x) {
        const double t = bf.spike_t[s];
        for (int k = 0; k < dm.K; ++k) {
            const svgpfa_latent_desc ds = bf.desc[k];
            const KConst kc = make_kconst(ds, bf.theta, bf.kscale, k);
            double mu = 0.0;
            for (int j = 0; j < ds.M; ++j) mu = fma(kappa_val(kc, t - zs[ds.moff + j]), as[ds.moff + j], mu);
            bf.mu_s[(size_t)s * dm.K + k] = mu;
        }
    }
}

// ------------------------------------------------------------------------------------------
// One warp per (trial, neuron) segment; lane <-> latent (K <= 32 per pass); 4 spike rows in flight.
constexpr int SG_THREADS = 256;

__global__ void __launch_bounds__(SG_THREADS) spike_gather_kernel(svgpfa_dims dm, svgpfa_buffers bf) {
    __shared__ double red[32];
    const int lane = threadIdx.x & 31;
    const int wpb = SG_THREADS / 32;
    const int64_t nseg = (int64_t)dm.R * dm.N;
    const int K = dm.K;
    double* gC = bf.shared + SVGPFA_SHARED_HDR;
    double val = 0.0;
    for (int64_t sg = (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); sg < nseg; sg += (int64_t)gridDim.x * wpb) {
        const int64_t s0 = bf.seg_off[sg], s1 = bf.seg_off[sg + 1];
        if (s0 == s1) continue;
        const int n = (int)(sg % dm.N);
        for (int k0 = 0; k0 < K; k0 += 32) {
            const int k = k0 + lane;
            if (k < K) {
                const double* p = bf.mu_s + (size_t)s0 * K + k;
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                int64_t s = s0;
                for (; s + 4 <= s1; s += 4) {
                    a0 += p[0];
                    a1 += p[K];
                    a2 += p[2 * (size_t)K];
                    a3 += p[3 * (size_t)K];
                    p += 4 * (size_t)K;
                }
                for (; s < s1; ++s) { a0 += *p; p += K; }
                const double sum = (a0 + a1) + (a2 + a3);
                atomicAdd(gC + (size_t)n * K + k, sum);
                val = fma(sum, bf.C[(size_t)n * K + k], val);
            }
        }
    }
    const double tot = block_sum(val, red);
    if (threadIdx.x == 0) atomicAdd(bf.shared + 4, tot);     // term2 (the d part is added by finalize)
}

}  // namespace

template <bool KGRAD, int UNROLL, int MINB>
static void launch_spike(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, cudaStream_t st, int nsm) {
    const int LG = (dims->KM + 31) / 32;
    const int gy = (LG + SP_WPB - 1) / SP_WPB;
    // enough warps to fill the machine a few times over: split every trial's neurons into chunks
    const long target_warps = (long)nsm * 64 * 4;
    long n_chunks = (target_warps + (long)dims->R * LG - 1) / ((long)dims->R * LG);
    if (n_chunks < 1) n_chunks = 1;
    if (n_chunks > dims->N) n_chunks = dims->N;
    const int chunk = (int)((dims->N + n_chunks - 1) / n_chunks);
    n_chunks = (dims->N + chunk - 1) / chunk;
    const dim3 grid((unsigned)(dims->R * n_chunks), gy);
    spike_fwd_bwd_kernel<KGRAD, UNROLL, MINB><<<grid, 32 * SP_WPB, 0, st>>>(*dims, *buf, flags, (int)n_chunks, chunk);
}

// SVGPFA_SPIKE_VARIANT (environment, experiments only) selects the kernel shape: UNROLL*10 + MINB.
static int spike_variant() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SVGPFA_SPIKE_VARIANT");
        v = e ? atoi(e) : 0;
    }
    return v;
}

extern "C" int svgpfa_spike_fwd_bwd(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, void* stream) {
    if (!dims || !buf) return svgpfa_set_error(SVGPFA_E_ARG, "spike_fwd_bwd", cudaSuccess);
    if (dims->R == 0 || dims->S == 0 || dims->N == 0) return SVGPFA_OK;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    cudaStream_t st = (cudaStream_t)stream;
    const bool kgrad = flags & (SVGPFA_GRAD_KERNEL | SVGPFA_GRAD_INDLOCS);
    const int var = spike_variant();
#define SP_CASE(code, U, MB)                                                     \
    case code:                                                                   \
        if (kgrad) launch_spike<true, U, MB>(dims, buf, flags, st, nsm);         \
        else launch_spike<false, U, MB>(dims, buf, flags, st, nsm);              \
        break;
    switch (var) {
        SP_CASE(24, 2, 4)
        SP_CASE(26, 2, 6)
        SP_CASE(28, 2, 8)
        SP_CASE(44, 4, 4)
        SP_CASE(46, 4, 6)
        SP_CASE(48, 4, 8)
        SP_CASE(84, 8, 4)
        default:
            if (kgrad) launch_spike<true, 4, 6>(dims, buf, flags, st, nsm);
            else launch_spike<false, 4, 6>(dims, buf, flags, st, nsm);
    }
#undef SP_CASE
    SVGPFA_CHECK_LAUNCH("spike_fwd_bwd");
    return SVGPFA_OK;
}

extern "C" int svgpfa_spike_latent_means(const svgpfa_dims* dims, const svgpfa_buffers* buf, void* stream) {
    if (!dims || !buf) return svgpfa_set_error(SVGPFA_E_ARG, "spike_latent_means", cudaSuccess);
    if (dims->R == 0 || dims->S == 0) return SVGPFA_OK;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    int n_split = (nsm * 8 + dims->R - 1) / dims->R;
    if (n_split < 1) n_split = 1;
    const size_t smem = sizeof(double) * 2 * (size_t)dims->KM;
    spike_means_kernel<<<dims->R * n_split, SM_THREADS, smem, (cudaStream_t)stream>>>(*dims, *buf, n_split);
    SVGPFA_CHECK_LAUNCH("spike_latent_means");
    return SVGPFA_OK;
}

int svgpfa_launch_spike_gather(const svgpfa_dims* dims, const svgpfa_buffers* buf, cudaStream_t stream) {
    if (dims->R == 0 || dims->S == 0) return SVGPFA_OK;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    const long nseg = (long)dims->R * dims->N;
    long blocks = (nseg + SG_THREADS / 32 - 1) / (SG_THREADS / 32);
    if (blocks > (long)nsm * 16) blocks = (long)nsm * 16;
    spike_gather_kernel<<<(unsigned)blocks, SG_THREADS, 0, stream>>>(*dims, *buf);
    SVGPFA_CHECK_LAUNCH("spike_gather");
    return SVGPFA_OK;
}

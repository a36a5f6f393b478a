// Spike-time (AssocTimes) term of the expected log-likelihood over ragged CSR spike trains:
//   spike_fwd_bwd_kernel     sum_s [ sum_k C[n_s,k] kappa_k(t_s, Z_kr) . alpha_kr ] and every adjoint in one pass
//   spike_means_kernel       mu_s[s][k] = kappa_k(t_s, Z_kr) . alpha_kr           (cached-statistics path)
//   spike_gather_kernel      sum_s sum_k mu_s[s][k] C[n_s,k], dC[n][k] += segment sums  (HBM-bound gather)
// Replaces stats/kernelsMatricesStore.py:208-221, stats/svPosteriorOnLatents.py:265-300 (mean only: the
// exponential link never uses the spike-time variance, expectedLogLikelihood.py:210-213),
// stats/svEmbedding.py:137-144.  Ktz[k][r] (S_r x M) is never materialised.
//
// Mapping of spike_tile_kernel: a warp owns 32 consecutive (latent, inducing point) pairs of one trial
// -- z_j, alpha_j and the kernel constants live in registers -- and walks the spikes of a range of neurons.
// Spikes are stored neuron-major inside a trial, so the embedding weight C[n,k] is constant over a segment:
//   pn_j(n)   = sum_{s in (r,n)} kappa(t_s - z_j)
//   abar_j    = sum_n C[n,k] pn_j(n)          (no cross-lane traffic)
//   dC[n,k]  += sum_j alpha_j pn_j(n)         (one segmented warp reduction per NON-EMPTY segment)
// and the value of the term is alpha . abar (taken in svgpfa_finalize).

#include "common.cuh"

namespace {

// segmented (by latent) sum over the lanes of a warp; valid in the first lane of every segment
__device__ __forceinline__ double seg_sum(double v, unsigned same_mask) {
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        const double o = __shfl_down_sync(0xffffffffu, v, 1 << b);
        if (same_mask & (1u << b)) v += o;
    }
    return v;
}

// Lane-local description of one (latent, inducing point) pair.
struct PairInfo {
    int li, k;
    bool active, head;
    unsigned same;
};

__device__ __forceinline__ PairInfo pair_info(const svgpfa_dims& dm, const svgpfa_buffers& bf, int li, int lane) {
    PairInfo p;
    p.li = li;
    p.active = li < dm.KM;
    const int l = p.active ? li : dm.KM - 1;
    int k = 0;
    while (k + 1 < dm.K && bf.desc[k + 1].moff <= l) ++k;
    p.k = k;
    // segment structure of the warp (lanes of one latent are contiguous)
    unsigned same = 0;
#pragma unroll
    for (int b = 0; b < 5; ++b) {
        const int ok = __shfl_down_sync(0xffffffffu, k, 1 << b);
        const int oact = __shfl_down_sync(0xffffffffu, (int)p.active, 1 << b);
        if (lane + (1 << b) < 32 && ok == k && oact && p.active) same |= 1u << b;
    }
    p.same = same;
    const int kprev = __shfl_up_sync(0xffffffffu, k, 1);
    p.head = p.active && (lane == 0 || kprev != k);
    return p;
}

// Lane slots of spike_tile_kernel: the pairs of the exponential-quadratic latents first, then -- starting at a CTA
// boundary -- those of the periodic latents, so that every CTA (whose warps share the spike tile and meet at its
// barriers) evaluates ONE kernel type: a periodic evaluation costs ~2.5x an exponential-quadratic one, and a warp whose
// lanes mix the types runs both loops one after the other.  Returns the storage index moff_k + j of the first of the
// np pairs of (warp group grp, lane), or a value >= KM for an idle lane.
__device__ __forceinline__ int slot_to_pair(const svgpfa_dims& dm, const svgpfa_buffers& bf, int grp, int lane, int np,
                                            int wpb) {
    int n_eq = 0;
    for (int k = 0; k < dm.K; ++k) n_eq += bf.desc[k].ktype == SVGPFA_KERNEL_PERIODIC ? 0 : bf.desc[k].M;
    const int eq_warps = ((n_eq + 32 * np - 1) / (32 * np) + wpb - 1) / wpb * wpb;
    const bool per = grp >= eq_warps;
    int rem = ((per ? grp - eq_warps : grp) * 32 + lane) * np;
    if (rem >= (per ? dm.KM - n_eq : n_eq)) return dm.KM;
    for (int k = 0; k < dm.K; ++k) {
        const svgpfa_latent_desc ds = bf.desc[k];
        if ((ds.ktype == SVGPFA_KERNEL_PERIODIC) != per) continue;
        if (rem < ds.M) return ds.moff + rem;
        rem -= ds.M;
    }
    return dm.KM;
}

// ------------------------------------------------------------------------------------------
// spike_tile_kernel: the spikes of the CTA's neuron range are staged through SHARED memory; the inner loop is
// 13 FP64 + ~5 other instructions per (spike, inducing point).  What the ncu captures of the first version (spike
// times and a 2048-entry exp table read with plain loads; removed) showed, and what this kernel does about it:
//   * its 2048-entry exp table cost ~6 shared-memory wavefronts per warp lookup (random 8-byte gathers; LSU data
//     pipe 80 % busy)  ->  svgpfa_exp2m: 256 entries x 16 replicas, conflict-free by construction (common.cuh);
//   * the warps of a CTA own 128 consecutive pairs of ONE trial and walk the same spikes, so a tile of ST_TILE spike
//     times is loaded once per CTA (coalesced), re-based to the tile's first spike (t' = t - t0: small magnitudes, so
//     the scaled difference below keeps full relative accuracy) and read with warp-uniform LDS; the segment ends
//     and the embedding weight are fetched one segment ahead;
//   * exponential-quadratic pairs use pre-scaled coordinates: w = t' sc - (z - t0) sc with
//     sc = sqrt(256 / (2 ln2)) / l, so that kappa = 2^(-w^2 / 256) and the moments are taken in w:
//     sum kappa, sum kappa w, sum kappa w^2, rescaled by 1/sc, 1/sc^2 once per lane;
//   * the range clamp of the exponent is decided per tile from the tile's [min, max] spike time (warp-uniform
//     branch to a clamping copy of the loop), not per evaluation.
// Where the time goes now (tools/probe_eval.py, tools/probe_issue.py): the evaluation sequence ALONE, without
// segments or staging, runs at 35.6 cycles per warp evaluation per SM sub-partition -- an FP64 instruction costs
// max(2, number of distinct register operands) issue cycles (DFMA with three register operands: 3.0-3.7) and the
// integer/LDS instructions are not hidden -- and the kernel reaches 44 cycles (81 % of that).  More pairs per lane
// (NP = 2, 4) amortise the per-segment work but were measured slower (register pressure, profiles/README.md).
// NP > 1 requires every M_k to be a multiple of NP (all pairs of a lane then share the latent).
// Register budget: __maxnreg__ instead of a min-blocks launch bound -- with a min-blocks hint ptxas assumes the
// resident warps hide latency and emits the evaluations of an iteration as one dependency chain after the other.
#ifndef ST_U1
#define ST_U1 8       // spikes per iteration with one pair per lane; measured on the 2000-trial shard: 2 -> 16.04 ms,
#endif                // 4 -> 14.43, 8 -> 14.05 (compile with -DST_U1=... to compare)
constexpr int ST_TILE = 1024;
constexpr int ST_MAX_WPB = 8;
constexpr int ST_SLOTS = 8;                                   // deferred dC reductions per warp
constexpr int ST_DCW = ST_SLOTS * 33 + ST_SLOTS / 2;          // doubles per warp: [slot][33] values + neuron ids
constexpr unsigned ST_SMEM = SVGPFA_EXP2M_TAB_BYTES + 8 * (ST_TILE + 2 * ST_MAX_WPB + ST_MAX_WPB * ST_DCW) + 16384;

// dC of up to ST_SLOTS finished segments of a single-latent warp: slot rows [slot][lane] hold alpha_j * sum kappa;
// lane (slot = lane & 7, part = lane >> 3) adds 8 entries of its row, two shuffles join the four parts, and lane
// `slot` issues the one atomic of the segment.  ~3 instructions per segment instead of a 5-step segmented warp
// reduction per segment (the flush was 14 % of the kernel's stall samples).
__device__ __forceinline__ void drain_dc(const double* __restrict__ dcb, const int* __restrict__ dcn, int cnt, int lane,
                                         double* __restrict__ gCk, int K) {
    __syncwarp();
    const int sl = lane & 7, part = lane >> 3;
    double sum = 0.0;
    const double* row = dcb + sl * 33 + 8 * part;
#pragma unroll
    for (int e = 0; e < 8; ++e) sum += row[e];
    sum += __shfl_xor_sync(0xffffffffu, sum, 8);
    sum += __shfl_xor_sync(0xffffffffu, sum, 16);
    if (part == 0 && sl < cnt) atomicAdd(gCk + (size_t)dcn[sl] * K, sum);
    __syncwarp();
}


// U spikes x NP pairs = NE independent evaluations, stage by stage (svgpfa_exp2m_n)
template <bool KGRAD, bool CLAMP, int U, int NP>
__device__ __forceinline__ void eq_eval_n(const double (&t)[U], double sc, const double (&zs)[NP], unsigned etab,
                                          double (&pn)[NP], double (&p1)[NP], double (&p2)[NP]) {
    constexpr int NE = U * NP;
    double w[NE], w2[NE], wc[NE], kv[NE];
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
        for (int p = 0; p < NP; ++p) w[u * NP + p] = fma(t[u], sc, zs[p]);
    svgpfa_pin(w);
#pragma unroll
    for (int e = 0; e < NE; ++e) w2[e] = w[e] * w[e];
    svgpfa_pin(w2);
#pragma unroll
    for (int e = 0; e < NE; ++e) wc[e] = CLAMP ? svgpfa_exp2m_clamp(w2[e]) : w2[e];
    // Full-accuracy variant (degree 4, <= 3.5e-16).  svgpfa_exp2m_n<NE, 3> (degree-3 economised polynomial, <= 1.8e-14,
    // + I2F range reduction) measures 35.6 -> 33.1 cycles per warp evaluation (tools/probe_eval.py) and passes every
    // per-evaluation parity test, but it is NOT used: with it the L-BFGS replay of the reference's own example
    // (tests/test_gpu_parity.py::test_config1_svem_replay[direct]) takes 9 instead of 10 iterations in one step -- a
    // termination test the reference passes by 1e-14.  This kernel is the exact-replay path; the speed comes from the
    // panel path (panel.cu).
    svgpfa_exp2m_n<NE>(wc, etab, kv);
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            const int e = u * NP + p;
            pn[p] += kv[e];
            if (KGRAD) {
                p1[p] = fma(kv[e], w[e], p1[p]);
                p2[p] = fma(kv[e], w2[e], p2[p]);
            }
        }
    }
}

// cnt spikes (warp-uniform shared-memory reads) against the NP pairs of the lane; U spikes per iteration, then
// remainders of 4, 2 and 1
template <bool KGRAD, bool CLAMP, int NP>
__device__ __forceinline__ void eq_run(const double* __restrict__ tp, int cnt, double sc, const double (&zs)[NP],
                                       unsigned etab, double (&pn)[NP], double (&p1)[NP], double (&p2)[NP]) {
    constexpr int U = NP == 1 ? ST_U1 : 4 / NP;
    const double* const pe = tp + (cnt - cnt % U);
#pragma unroll 1
    for (; tp != pe; tp += U) {
        double t[U];
#pragma unroll
        for (int u = 0; u < U; ++u) t[u] = tp[u];
        eq_eval_n<KGRAD, CLAMP, U, NP>(t, sc, zs, etab, pn, p1, p2);
    }
    if (U == 8 && (cnt & 4)) {
        const double t[4] = {tp[0], tp[1], tp[2], tp[3]};
        eq_eval_n<KGRAD, CLAMP, 4, NP>(t, sc, zs, etab, pn, p1, p2);
        tp += 4;
    }
    if (U >= 4 && (cnt & 2)) {
        const double t[2] = {tp[0], tp[1]};
        eq_eval_n<KGRAD, CLAMP, 2, NP>(t, sc, zs, etab, pn, p1, p2);
        tp += 2;
    }
    if (U >= 2 && (cnt & 1)) {
        const double t[1] = {tp[0]};
        eq_eval_n<KGRAD, CLAMP, 1, NP>(t, sc, zs, etab, pn, p1, p2);
    }
}

// periodic pairs: w = sin(pi d/p) sc with sc = sqrt(256 * 2 / ln2) / l;  p1 += kappa sin(2 pi d/p),
// p2 += kappa w^2 (rescaled by 1/sc^2 at the end), p3 += kappa sin(2 pi d/p) d
// sin and cos of 2 pi x by table: n = rint(64 x), u = 64 x - n in [-1/2, 1/2], theta = 2 pi u / 64 (|theta| <= 0.0491),
//   sin(2 pi x) = S_n cos(theta) + C_n sin(theta),  cos(2 pi x) = C_n cos(theta) - S_n sin(theta)
// with (S_n, C_n) = sincos(2 pi n / 64) in shared memory (64 entries x 16 replicas, conflict-free like the exp table) and
// Taylor polynomials of degree 7 / 8 in theta (truncation < 5e-18).  19 FP64 instructions against ~40 (+ ~20 others) of
// libdevice's sincospi, which computes a full-range sine AND cosine polynomial.
constexpr int ST_SC_ENTRIES = SVGPFA_SC_ENTRIES;
constexpr int ST_SC_DOUBLES = 2 * ST_SC_ENTRIES * SVGPFA_EXP2M_REP;        // 16 KB

__device__ __forceinline__ void load_sincos_tab(double2* tab) { svgpfa_load_sincos_tab<SVGPFA_EXP2M_REP>(tab); }

__device__ __forceinline__ void sincos2pi_tab(double x64, const double2* __restrict__ lane_sc, double& sv, double& cv) {
    svgpfa_sincos2pi_tab<SVGPFA_EXP2M_REP>(x64, lane_sc, sv, cv);
}

// kappa = s2 exp(nh sin^2(pi d/p)) with sin^2 = (1 - cos(2 pi d/p)) / 2;  hsc2 = sc^2 / 2, invp64 = 64 / p
template <bool KGRAD>
__device__ __forceinline__ void per_eval(double t, double hsc2, double zc, double invp64, unsigned etab,
                                         const double2* __restrict__ lane_sc, double& pn, double& p1, double& p2,
                                         double& p3) {
    const double dl = t - zc;
    double s2x, c2x;
    sincos2pi_tab(dl * invp64, lane_sc, s2x, c2x);
    const double w2 = fabs((1.0 - c2x) * hsc2);               // |.|: a rounding-level negative value must not look huge to the clamp
    const double kv = svgpfa_exp2m(svgpfa_exp2m_clamp(w2), etab);
    pn += kv;
    if (KGRAD) {
        const double ww = kv * s2x;
        p1 += ww;
        p2 = fma(kv, w2, p2);
        p3 = fma(ww, dl, p3);
    }
}

// two spikes per iteration: the sincos + exp chain is ~35 dependent FP64 instructions, so a second independent
// chain per lane roughly halves the exposed latency (config #3 ran at ~400 cycles per periodic warp evaluation with
// libdevice's sincospi and one spike per iteration)
template <bool KGRAD, int NP>
__device__ __forceinline__ void per_run(const double* __restrict__ tp, int cnt, double sc, const double (&zc)[NP],
                                        double invp, unsigned etab, const double2* __restrict__ lane_sc,
                                        double (&pn)[NP], double (&p1)[NP], double (&p2)[NP], double (&p3)[NP]) {
    const double hsc2 = 0.5 * sc * sc, invp64 = invp * ST_SC_ENTRIES;
    int i = 0;
#pragma unroll 1
    for (; i + 2 <= cnt; i += 2) {
        const double t0 = tp[i], t1 = tp[i + 1];
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            double qn = 0.0, q1 = 0.0, q2 = 0.0, q3 = 0.0;               // second chain, joined below
            per_eval<KGRAD>(t0, hsc2, zc[p], invp64, etab, lane_sc, pn[p], p1[p], p2[p], p3[p]);
            per_eval<KGRAD>(t1, hsc2, zc[p], invp64, etab, lane_sc, qn, q1, q2, q3);
            pn[p] += qn; p1[p] += q1; p2[p] += q2; p3[p] += q3;
        }
    }
    if (i < cnt) {
#pragma unroll
        for (int p = 0; p < NP; ++p) per_eval<KGRAD>(tp[i], hsc2, zc[p], invp64, etab, lane_sc, pn[p], p1[p], p2[p], p3[p]);
    }
}

template <bool KGRAD, int NP, int MAXT, int MAXR>
__global__ void __launch_bounds__(MAXT) __maxnreg__(MAXR) spike_tile_kernel(svgpfa_dims dm, svgpfa_buffers bf, uint32_t flags,
                                                                int n_chunks, int chunk, int has_periodic) {
    extern __shared__ __align__(16) double st_smem[];       // ST_SMEM bytes: table | spike tile | min/max scratch
    double* ts = st_smem + SVGPFA_EXP2M_TAB_BYTES / 8;
    double* red = ts + ST_TILE;
    double* dcb = red + 2 * ST_MAX_WPB + (threadIdx.x >> 5) * ST_DCW;          // this warp's deferred-dC rows
    int* dcn = reinterpret_cast<int*>(dcb + ST_SLOTS * 33);
    // sincos table of the periodic kernels, after the per-warp rows (allocated only when a latent is periodic)
    double2* sctab = reinterpret_cast<double2*>(red + 2 * ST_MAX_WPB + (blockDim.x >> 5) * ST_DCW);
    if (has_periodic) load_sincos_tab(sctab);
    const double2* lane_sc = sctab + (threadIdx.x & (SVGPFA_EXP2M_REP - 1));
    svgpfa_load_exp2m_tab(st_smem);
    const unsigned etab = svgpfa_exp2m_lane_tab(st_smem);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int rl = blockIdx.x / n_chunks, r = dm.r0 + rl, nc = blockIdx.x - rl * n_chunks;
    const int grp = blockIdx.y * wpb + warp;
    // a warp past the last pair still takes part in the staging and the barriers; its lanes are inactive
    const PairInfo pi = pair_info(dm, bf, slot_to_pair(dm, bf, grp, lane, NP, wpb), lane);   // first pair of the lane
    const svgpfa_latent_desc ds = bf.desc[pi.k];
    const KConst kc = make_kconst(ds, bf.theta, bf.kscale, pi.k);
    const bool per = kc.type == SVGPFA_KERNEL_PERIODIC;
    const double sc = sqrt(-kc.nh * SVGPFA_EXP2M_INV_L);
    double z[NP], a[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        const int l = pi.active ? pi.li + p : dm.KM - 1;
        z[p] = bf.Z[(size_t)dm.R * ds.moff + (size_t)r * ds.M + (l - ds.moff)];
        a[p] = pi.active ? kc.s2 * bf.alpha[(size_t)r * dm.KM + l] : 0.0;     // scale^2 alpha_j
    }
    const bool need_emb = flags & SVGPFA_GRAD_EMBEDDING;
    const bool warp_live = __any_sync(0xffffffffu, pi.active);
    // all 32 lanes active and on the same latent (M_k a multiple of 32 / NP): dC reductions are deferred
    const int k_lane0 = __shfl_sync(0xffffffffu, pi.k, 0);       // (not inside the && below: every lane must shuffle)
    const bool one_latent = __all_sync(0xffffffffu, pi.active && pi.k == k_lane0);
    int slot = 0;
    const int nb = nc * chunk, ne = min(dm.N, nb + chunk);
    const int64_t* __restrict__ seg = bf.seg_off + (size_t)r * dm.N;
    const double* __restrict__ st = bf.spike_t;
    const double* __restrict__ Ck = bf.C + pi.k;
    double* gC = bf.shared + SVGPFA_SHARED_HDR + pi.k;
    const int64_t S0 = seg[nb], S1 = seg[ne];
    double abar[NP], dz[NP], d0[NP], d1[NP], pn[NP], p1[NP], p2[NP], p3[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) abar[p] = dz[p] = d0[p] = d1[p] = pn[p] = p1[p] = p2[p] = p3[p] = 0.0;
    // segment walk: n = current neuron, seg_end = end of its segment, seg_nxt = end of the next one (loaded one
    // segment ahead so that its latency is hidden behind a segment's worth of arithmetic)
    int n = nb;
    int64_t seg_end = seg[nb + 1], pos = S0;
    int64_t seg_nxt = seg[min(nb + 2, ne)];
    double c = Ck[(size_t)n * dm.K];
    for (int64_t tile0 = S0; tile0 < S1; tile0 += ST_TILE) {
        const int len = (int)min((int64_t)ST_TILE, S1 - tile0);
        __syncthreads();                       // the previous tile has been consumed (and the table is loaded)
        const double t0 = st[tile0];
        double lo = 0.0, hi = 0.0;             // t' of the tile's first spike is 0
        for (int i = threadIdx.x; i < len; i += blockDim.x) {
            const double v = st[tile0 + i] - t0;
            ts[i] = v;
            lo = fmin(lo, v);
            hi = fmax(hi, v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0) { red[warp] = lo; red[ST_MAX_WPB + warp] = hi; }
        __syncthreads();
        if (!warp_live) continue;
        for (int w = 0; w < wpb; ++w) { lo = fmin(lo, red[w]); hi = fmax(hi, red[ST_MAX_WPB + w]); }
        double zc[NP], zs[NP];
        bool over = false;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            zc[p] = z[p] - t0;
            zs[p] = -zc[p] * sc;
            const double wm = fmax(fabs(fma(lo, sc, zs[p])), fabs(fma(hi, sc, zs[p])));
            over |= !(wm * wm < SVGPFA_EXP2M_LIMIT);
        }
        const bool clamp = __any_sync(0xffffffffu, !per && over);
        const int64_t tile1 = tile0 + len;
        while (pos < tile1) {
            while (seg_end <= pos) {                                       // next non-empty segment (pos < S1 here)
                ++n;
                seg_end = seg_nxt;
                seg_nxt = seg[min(n + 2, ne)];
                c = Ck[(size_t)n * dm.K];
            }
            const int64_t e = min(seg_end, tile1);
            const int cnt = (int)(e - pos);
            const double* tp = ts + (int)(pos - tile0);
            if (per) per_run<KGRAD, NP>(tp, cnt, sc, zc, kc.invp, etab, lane_sc, pn, p1, p2, p3);
            else if (clamp) eq_run<KGRAD, true, NP>(tp, cnt, sc, zs, etab, pn, p1, p2);
            else eq_run<KGRAD, false, NP>(tp, cnt, sc, zs, etab, pn, p1, p2);
            __syncwarp();
            pos = e;
            if (pos == seg_end) {                                          // segment (r, n) complete
                double v = 0.0;
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    abar[p] = fma(c, pn[p], abar[p]);
                    if (KGRAD) {
                        dz[p] = fma(c, p1[p], dz[p]);
                        d0[p] = fma(c, p2[p], d0[p]);
                        d1[p] = fma(c, p3[p], d1[p]);
                    }
                    v = fma(pn[p], a[p], v);
                    pn[p] = p1[p] = p2[p] = p3[p] = 0.0;
                }
                if (need_emb) {
                    if (one_latent) {
                        dcb[slot * 33 + lane] = v;
                        if (lane == 0) dcn[slot] = n;
                        if (++slot == ST_SLOTS) {
                            drain_dc(dcb, dcn, ST_SLOTS, lane, gC, dm.K);
                            slot = 0;
                        }
                    } else {
                        v = seg_sum(v, pi.same);
                        if (pi.head) atomicAdd(gC + (size_t)n * dm.K, v);
                    }
                }
            }
        }
    }
    if (!warp_live) return;
    if (slot) drain_dc(dcb, dcn, slot, lane, gC, dm.K);
    const double isc = 1.0 / sc;
    double t0s = 0.0, t1s = 0.0;
#pragma unroll
    for (int p = 0; p < NP; ++p) {
        if (pi.active) {
            atomicAdd(bf.abar_spk + (size_t)r * dm.KM + pi.li + p, kc.s2 * abar[p]);
            // kbar_sj = C alpha_j ; d delta/dz = -1 ; dkappa/ddelta = kappa * (delta | sin 2 pi d/p) * dd
            if (KGRAD && (flags & SVGPFA_GRAD_INDLOCS))
                atomicAdd(bf.dz_acc + (size_t)r * dm.KM + pi.li + p, -a[p] * kc.dd * (per ? dz[p] : dz[p] * isc));
            // dkappa/dtheta0 = kappa (d^2 | sin^2) dl;  periodic: dkappa/dtheta1 = kappa sin(2 pi d/p) d dp
            t0s = fma(a[p] * kc.dl, d0[p] * isc * isc, t0s);
            t1s = fma(a[p] * kc.dp, d1[p], t1s);
        }
    }
    if (KGRAD && (flags & SVGPFA_GRAD_KERNEL)) {
        const double t0 = seg_sum(t0s, pi.same);
        const double t1 = seg_sum(t1s, pi.same);
        if (pi.head) {
            double* dth = bf.dth_part + (size_t)r * dm.TH + ds.thoff;
            atomicAdd(dth, t0);
            if (ds.nth > 1) atomicAdd(dth + 1, t1);
        }
    }
}

// ------------------------------------------------------------------------------------------
constexpr int SM_THREADS = 256;

__global__ void __launch_bounds__(SM_THREADS) spike_means_kernel(svgpfa_dims dm, svgpfa_buffers bf, int n_split) {
    extern __shared__ __align__(16) double sm[];
    double* etab = sm;                                   // replicated exp table (svgpfa_exp2m)
    double* zs = etab + SVGPFA_EXP2M_TAB_BYTES / 8;      // KM   z_j sc_k (pre-scaled; periodic: z_j)
    double* as = zs + dm.KM;                             // KM   scale^2 alpha
    double* ksc = as + dm.KM;                            // K    sc_k
    double* kip = ksc + dm.K;                            // K    1/p (0 for the exponential-quadratic kernel)
    svgpfa_load_exp2m_tab(etab);
    const unsigned lane_tab = svgpfa_exp2m_lane_tab(etab);
    const int rl = blockIdx.x / n_split, r = dm.r0 + rl, part = blockIdx.x - rl * n_split;
    for (int k = 0; k < dm.K; ++k) {
        const svgpfa_latent_desc ds = bf.desc[k];
        const KConst kc = make_kconst(ds, bf.theta, bf.kscale, k);
        const double sc = sqrt(-kc.nh * SVGPFA_EXP2M_INV_L);
        const bool per = kc.type == SVGPFA_KERNEL_PERIODIC;
        if (threadIdx.x == 0) { ksc[k] = sc; kip[k] = per ? kc.invp : 0.0; }
        for (int j = threadIdx.x; j < ds.M; j += blockDim.x) {
            const double z = bf.Z[(size_t)dm.R * ds.moff + (size_t)r * ds.M + j];
            zs[ds.moff + j] = per ? z : z * sc;
            as[ds.moff + j] = kc.s2 * bf.alpha[(size_t)r * dm.KM + ds.moff + j];
        }
    }
    __syncthreads();
    const int64_t s0 = bf.seg_off[(size_t)r * dm.N], s1 = bf.seg_off[(size_t)(r + 1) * dm.N];
    for (int64_t s = s0 + (int64_t)part * blockDim.x + threadIdx.x; s < s1; s += (int64_t)n_split * blockDim.x) {
        const double t = bf.spike_t[s];
        for (int k = 0; k < dm.K; ++k) {
            const int M = bf.desc[k].M, off = bf.desc[k].moff;
            const double sc = ksc[k], ip = kip[k];
            double mu = 0.0;
            if (ip == 0.0) {
                const double tsc = t * sc;
#pragma unroll 4
                for (int j = 0; j < M; ++j) {
                    const double w = tsc - zs[off + j];
                    mu = fma(svgpfa_exp2m(svgpfa_exp2m_clamp(w * w), lane_tab), as[off + j], mu);
                }
            } else {
#pragma unroll 2
                for (int j = 0; j < M; ++j) {
                    const double w = sinpi((t - zs[off + j]) * ip) * sc;
                    mu = fma(svgpfa_exp2m(svgpfa_exp2m_clamp(w * w), lane_tab), as[off + j], mu);
                }
            }
            bf.mu_s[(size_t)s * dm.K + k] = mu;
        }
    }
}

// ------------------------------------------------------------------------------------------
// Ragged gather of the cached-statistics path: gsum[n][k] += sum_{s in (r,n)} mu_s[s][k] over the shard's trials.
// The rows of one (trial, neuron) segment are contiguous in mu_s ([s][k]), so a segment is a contiguous block of
// cnt*K doubles.  One warp per segment; lane = (row, latent) with 32 / K rows per pass (K = 20: 20 lanes, K = 3:
// 30 lanes -- the round-1 mapping lane <-> latent left 3 of 32 lanes busy there), four passes in flight; the lanes of
// one latent are joined through shuffles at the end of the segment.  HBM-bound: S*K*8 bytes are read once.
// With the statistics cached the spike part of the expected log-likelihood is LINEAR in C,
//     sum_s sum_k mu_s[s][k] C[n_s][k] = sum_{n,k} gsum[n][k] C[n][k],
// so the gather runs once per embedding M-step (svgpfa_cached_ell_fwd_bwd with SVGPFA_REUSE_SPIKE skips it) and every
// closure evaluation after the first costs N*K multiply-adds for this term.
constexpr int SG_THREADS = 256;

__global__ void __launch_bounds__(SG_THREADS) spike_gather_kernel(svgpfa_dims dm, svgpfa_buffers bf) {
    const int lane = threadIdx.x & 31;
    const int wpb = SG_THREADS / 32;
    const int64_t seg0 = (int64_t)dm.r0 * dm.N, nseg = seg0 + (int64_t)(dm.rn ? dm.rn : dm.R) * dm.N;
    const int K = dm.K;
    const int kp = K < 32 ? K : 32;                        // latents per pass
    const int rows = 32 / kp;                              // spike rows per pass
    const int row = lane / kp, kl = lane - row * kp;
    const bool live = row < rows;
    for (int64_t sg = seg0 + (int64_t)blockIdx.x * wpb + (threadIdx.x >> 5); sg < nseg; sg += (int64_t)gridDim.x * wpb) {
        const int64_t s0 = bf.seg_off[sg], s1 = bf.seg_off[sg + 1];
        if (s0 == s1) continue;
        const int n = (int)(sg % dm.N);
        for (int k0 = 0; k0 < K; k0 += 32) {
            const int k = k0 + kl;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            if (live && k < K) {
                const double* p = bf.mu_s + (size_t)(s0 + row) * K + k;
                const size_t step = (size_t)rows * K;
                int64_t s = s0 + row;
                for (; s + 3 * rows < s1; s += 4 * rows) {
                    a0 += p[0];
                    a1 += p[step];
                    a2 += p[2 * step];
                    a3 += p[3 * step];
                    p += 4 * step;
                }
                for (; s < s1; s += rows) { a0 += *p; p += step; }
            }
            double sum = (a0 + a1) + (a2 + a3);
            for (int rr = 1; rr < rows; ++rr) {            // join the rows of one latent: lane kl collects lanes kl + rr*kp
                const double o = __shfl_sync(0xffffffffu, sum, (kl + rr * kp) & 31);
                if (row == 0) sum += o;
            }
            if (row == 0 && k < K) atomicAdd(bf.gsum + (size_t)n * K + k, sum);
        }
    }
}

// shared[4] += sum_{n,k} gsum[n][k] C[n][k];  dC += gsum          (one block; N*K is small)
__global__ void __launch_bounds__(256) gsum_apply_kernel(svgpfa_dims dm, svgpfa_buffers bf) {
    __shared__ double red[32];
    double* gC = bf.shared + SVGPFA_SHARED_HDR;
    double val = 0.0;
    const int NK = dm.N * dm.K;
    for (int i = threadIdx.x; i < NK; i += blockDim.x) {
        const double g = bf.gsum[i];
        val = fma(g, bf.C[i], val);
        gC[i] += g;                                        // the quadrature kernel's atomics on dC are complete (stream order)
    }
    const double tot = block_sum(val, red);
    if (threadIdx.x == 0) bf.shared[4] += tot;             // term2 (the d part is added by finalize)
}

}  // namespace

// Grid of the spike kernels: warps of 32 * np pairs, wpb warps per CTA (the value in [min_wpb, max_wpb] that leaves
// the fewest idle warps), and enough CTAs to fill the machine a few times over (every trial's neurons are split into
// n_chunks ranges when R alone does not provide them).
static void spike_grid(const svgpfa_dims* dims, int nsm, int np, int min_wpb, int max_wpb, bool by_type, dim3* grid,
                       int* wpb_out, int* n_chunks_out, int* chunk_out) {
    // warp groups: by_type (spike_tile_kernel) = those of the exponential-quadratic pairs, rounded up to whole CTAs,
    // then those of the periodic pairs (slot_to_pair); otherwise ceil(KM / (32 np))
    int n_eq = 0;
    for (int k = 0; k < dims->K; ++k)
        n_eq += (by_type && dims->desc_host[k].ktype == SVGPFA_KERNEL_PERIODIC) ? 0 : dims->desc_host[k].M;
    const int lg_eq = (n_eq + 32 * np - 1) / (32 * np), lg_per = (dims->KM - n_eq + 32 * np - 1) / (32 * np);
    auto warps_for = [&](int w) { return (lg_eq + w - 1) / w * w + (lg_per + w - 1) / w * w; };
    const int LG = lg_eq + lg_per;
    int wpb = LG < min_wpb ? (LG > 0 ? LG : 1) : min_wpb, best = 1 << 30;
    if (lg_eq && lg_per && wpb < min_wpb) wpb = lg_eq > lg_per ? lg_eq : lg_per;
    for (int w = min_wpb; w <= max_wpb && LG >= min_wpb; ++w) {
        const int waste = warps_for(w) - LG;
        if (waste < best) { best = waste; wpb = w; }
    }
    const int gy = warps_for(wpb) / wpb;
    const long target_warps = (long)nsm * 64 * 4 / np;
    const long Rn = svgpfa_ntrials(dims);
    long n_chunks = (target_warps + Rn * LG - 1) / (Rn * LG);
    if (dims->spike_chunks > 0) n_chunks = dims->spike_chunks;                 // tests: force long neuron ranges
    if (n_chunks < 1) n_chunks = 1;
    if (n_chunks > dims->N) n_chunks = dims->N;
    const int chunk = (int)((dims->N + n_chunks - 1) / n_chunks);
    n_chunks = (dims->N + chunk - 1) / chunk;
    *grid = dim3((unsigned)(Rn * n_chunks), gy);
    *wpb_out = wpb;
    *n_chunks_out = (int)n_chunks;
    *chunk_out = chunk;
}

// One pair per lane, 128 threads, 128 registers (4 CTAs per SM).  Measured on B200, config #5 shard of 2000 trials
// (ms): the first register/global-load kernel 17.16, 96-register build 15.92, this one 15.22 -> 14.05, 2 pairs per
// lane 16.43, 4 pairs per lane 18.4 (profiles/README.md; the slower variants are no longer built).
template <int NP, int MAXT, int MAXR>
static void launch_tile(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, cudaStream_t st, int nsm,
                        bool kgrad) {
    SVGPFA_ENSURE_SMEM(ST_SMEM, spike_tile_kernel<true, NP, MAXT, MAXR>);
    SVGPFA_ENSURE_SMEM(ST_SMEM, spike_tile_kernel<false, NP, MAXT, MAXR>);
    dim3 grid;
    int wpb, n_chunks, chunk;
    spike_grid(dims, nsm, NP, MAXT >= 256 ? 4 : (MAXT / 32 < 4 ? MAXT / 32 : 4), MAXT / 32, true, &grid, &wpb, &n_chunks, &chunk);
    int has_periodic = 0;
    for (int k = 0; k < dims->K; ++k) has_periodic |= dims->desc_host[k].ktype == SVGPFA_KERNEL_PERIODIC;
    const size_t smem = SVGPFA_EXP2M_TAB_BYTES + 8 * (size_t)(ST_TILE + 2 * ST_MAX_WPB + wpb * ST_DCW) +
                        (has_periodic ? 16384 : 0);
    if (kgrad) spike_tile_kernel<true, NP, MAXT, MAXR><<<grid, 32 * wpb, smem, st>>>(*dims, *buf, flags, n_chunks, chunk, has_periodic);
    else spike_tile_kernel<false, NP, MAXT, MAXR><<<grid, 32 * wpb, smem, st>>>(*dims, *buf, flags, n_chunks, chunk, has_periodic);
}

extern "C" int svgpfa_spike_fwd_bwd(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, void* stream) {
    if (!dims || !buf) return svgpfa_set_error(SVGPFA_E_ARG, "spike_fwd_bwd", cudaSuccess);
    if (dims->R == 0 || dims->S == 0 || dims->N == 0) return SVGPFA_OK;
    if (!dims->desc_host) return svgpfa_set_error(SVGPFA_E_ARG, "spike_fwd_bwd: dims.desc_host", cudaSuccess);
    const bool kgrad = flags & (SVGPFA_GRAD_KERNEL | SVGPFA_GRAD_INDLOCS);
    launch_tile<1, 128, 128>(dims, buf, flags, (cudaStream_t)stream, svgpfa_sm_count(), kgrad);
    SVGPFA_CHECK_LAUNCH("spike_fwd_bwd");
    return SVGPFA_OK;
}

extern "C" int svgpfa_spike_latent_means(const svgpfa_dims* dims, const svgpfa_buffers* buf, void* stream) {
    if (!dims || !buf) return svgpfa_set_error(SVGPFA_E_ARG, "spike_latent_means", cudaSuccess);
    if (dims->R == 0 || dims->S == 0) return SVGPFA_OK;
    const int nsm = svgpfa_sm_count();
    const int Rn = svgpfa_ntrials(dims);
    int n_split = (nsm * 8 + Rn - 1) / Rn;
    if (n_split < 1) n_split = 1;
    const size_t smem = SVGPFA_EXP2M_TAB_BYTES + sizeof(double) * (2 * (size_t)dims->KM + 2 * (size_t)dims->K);
    SVGPFA_ENSURE_SMEM(smem, spike_means_kernel);
    spike_means_kernel<<<Rn * n_split, SM_THREADS, smem, (cudaStream_t)stream>>>(*dims, *buf, n_split);
    SVGPFA_CHECK_LAUNCH("spike_latent_means");
    return SVGPFA_OK;
}

int svgpfa_launch_spike_gather(const svgpfa_dims* dims, const svgpfa_buffers* buf, bool reuse, cudaStream_t stream) {
    if (!buf->gsum) return svgpfa_set_error(SVGPFA_E_ARG, "cached_ell_fwd_bwd: gsum", cudaSuccess);
    if (dims->N == 0) return SVGPFA_OK;
    if (!reuse) {
        cudaMemsetAsync(buf->gsum, 0, sizeof(double) * (size_t)dims->N * dims->K, stream);
        if (dims->R > 0 && dims->S > 0) {
            const int nsm = svgpfa_sm_count();
            const long nseg = (long)svgpfa_ntrials(dims) * dims->N;
            long blocks = (nseg + SG_THREADS / 32 - 1) / (SG_THREADS / 32);
            if (blocks > (long)nsm * 16) blocks = (long)nsm * 16;
            spike_gather_kernel<<<(unsigned)blocks, SG_THREADS, 0, stream>>>(*dims, *buf);
            SVGPFA_CHECK_LAUNCH("spike_gather");
        }
    }
    gsum_apply_kernel<<<1, 256, 0, stream>>>(*dims, *buf);
    SVGPFA_CHECK_LAUNCH("gsum_apply");
    return SVGPFA_OK;
}

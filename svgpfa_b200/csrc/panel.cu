// Panel path of the spike-time (AssocTimes) term: the same quantities as spike.cu -- abar_spk, dC, dz_acc, dth_part,
// reference stats/kernelsMatricesStore.py:208-221, stats/svPosteriorOnLatents.py:265-300, stats/svEmbedding.py:137-144,
// stats/expectedLogLikelihood.py:210-213 -- without touching a spike per evaluation.
//
// Everything the term needs from the spikes of (trial r, neuron n) is a sum  sum_s f(t_s)  of functions f in the span
// of kappa_k(. - z_j) and of their derivatives with respect to z_j and theta_k: analytic functions of t whose scale of
// variation is the kernel length scale.  Cut the time axis into B panels; on a panel, replace f by its interpolant at
// P = 16 first-kind Chebyshev nodes t_i:  f(t) ~ sum_i l_i(t) f(t_i)  (l_i = Lagrange cardinal functions).  Then
//     sum_{s in (r,n)} f(t_s) = sum_i tau[r][n][i] f(t_i),      tau[r][n][i] = sum_{s in (r,n)} l_i(t_s),
// and tau -- NB = 16 B numbers per (trial, neuron) -- is STATIC: it depends on the data and the panelisation only.
// With c_s = C[n_s, k]:
//     mt[r][k][i]  = sum_n C[n,k] tau[r][n][i]                                    (skinny GEMM over the neurons)
//     abar_spk_j   = sum_s c_s kappa(t_s - z_j)       = sum_i mt_i kappa(t_i - z_j)
//     dz_j, dtheta = the same sums with dkappa/dz_j, dkappa/dtheta
//     mun[r][k][i] = sum_j alpha_j kappa(t_i - z_j)                               (latent mean at the nodes)
//     dC[n,k]     += sum_{s in (r,n)} mu_k(t_s)       = sum_i tau[r][n][i] mun[r][k][i]     (skinny GEMM over the nodes)
// i.e. the spikes of a trial are replaced by NB weighted pseudo-spikes shared by all neurons.  Work per evaluation:
// R NB (2 K M kernel values + 2 N K multiply-adds) against S K M kernel values (config #5: NB = 128 against ~10^4
// spikes per trial).  Accuracy (tests/test_gpu_panel.py, tools/panel_accuracy.py): relative error <= 3e-14 of every
// sum for panel half-width <= 0.625 length scale (exponential-quadratic) -- the caller chooses B from the current
// hyper-parameters (svgpfa_b200/model.py) and uses the direct kernel when B would exceed 32.
#include "common.cuh"

namespace {

constexpr int PM_P = SVGPFA_PM_P;

__device__ __forceinline__ void pm_dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// time of node i (panel i / P, Chebyshev node i % P)
__device__ __forceinline__ double pm_node_time(const svgpfa_dims& dm, int i) {
    const int b = i / PM_P, ii = i - b * PM_P;
    return dm.pm_lo + dm.pm_w * (b + 0.5 + 0.5 * cospi((ii + 0.5) / PM_P));
}

// ------------------------------------------------------------------------------------------
// tau[r][n][b P + i] = sum_{s in (r,n), t_s in panel b} l_i(x_s),  l_i(x) = sum_m D[i][m] T_m(x),
// D[i][m] = (2/P) w_m T_m(x_i)  (w_0 = 1/2; first-kind Chebyshev nodes x_i).
// Warp per (trial, neuron) segment, in two steps on the FP64 tensor path:
//   1. Chebyshev moments of the segment's panels,  c[m][b] = sum_{s in panel b} T_m(x_s):  lane <-> spike (32 per
//      iteration, one 14-step recurrence each, no repetition), the T values go through a shared-memory tile into
//      A fragments and are summed per panel by mma against the INDICATOR matrix  Ind[s][b] = (panel(s) == b)  --
//      a segmented reduction with exact FP64 accumulation and no atomics / read-modify-writes;
//   2. tau = D c, a 16 x 16 by 16 x B product per segment (8 mma per 8 panels), written straight from the fragments.
// Round-2 history: the first version gave 8 lanes to a spike, each repeating the recurrence and summing 16 terms for its
// node pair into shared-memory accumulators (13 ms per 20 000 trials of config #5: 240 FP64 lane-operations per spike and
// a zero / read-modify-write / reduce cycle of the accumulator rows per segment).  Rebuilt per block by the
// host-buffer entry whenever new spikes arrive.
// ------------------------------------------------------------------------------------------
constexpr int PMK_WARPS = 4;
constexpr int PMK_LDT = 20;                              // leading dimension of the [spike][degree] tile (conflict-free fragments)

template <int NT>                                         // NT = ceil(panels / 8) column tiles
__global__ void __launch_bounds__(32 * PMK_WARPS) panel_moments_kernel(svgpfa_dims dm, svgpfa_buffers bf, int n_chunks, int chunk) {
    constexpr int LDC = 8 * NT + 4;
    __shared__ double Ts_all[PMK_WARPS][32 * PMK_LDT];
    __shared__ double Cs_all[PMK_WARPS][PM_P * LDC];
    const int NB = dm.pm_B * PM_P;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, tg = lane & 3;
    double* Ts = Ts_all[warp];
    double* Cs = Cs_all[warp];
    double dfrag[2][4];                                   // A fragments of D: D[8 it + g][4 ks + tg]
#pragma unroll
    for (int it = 0; it < 2; ++it)
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
            const int i = 8 * it + g, m = 4 * ks + tg;
            dfrag[it][ks] = (m == 0 ? 1.0 : 2.0) / PM_P * cospi(m * (i + 0.5) / PM_P);
        }
    const int rl = blockIdx.x / n_chunks, r = dm.r0 + rl, nc = blockIdx.x - rl * n_chunks;
    const int nb = nc * chunk, ne = min(dm.N, nb + chunk);
    const double inv_w = 1.0 / dm.pm_w;
    for (int n = nb + warp; n < ne; n += PMK_WARPS) {
        const int64_t s0 = bf.seg_off[(size_t)r * dm.N + n], s1 = bf.seg_off[(size_t)r * dm.N + n + 1];
        double* out = bf.pm_tau + ((size_t)r * dm.N + n) * NB;
        if (s1 <= s0) {                                   // no spikes: a row of zeros
            for (int e = lane; e < NB; e += 32) out[e] = 0.0;
            continue;
        }
        double c[2][NT][2];
#pragma unroll
        for (int it = 0; it < 2; ++it)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt) c[it][nt][0] = c[it][nt][1] = 0.0;
        for (int64_t base = s0; base < s1; base += 32) {
            const int64_t s = base + lane;
            int pb = -1;                                  // panel of this lane's spike; -1: no spike (matches no column)
            double t[PM_P];
            if (s < s1) {
                const double rel = (bf.spike_t[s] - dm.pm_lo) * inv_w;
                pb = max(0, min(dm.pm_B - 1, (int)floor(rel)));
                const double x = fma(2.0, rel - (double)pb, -1.0), x2 = 2.0 * x;
                t[0] = 1.0;
                t[1] = x;
#pragma unroll
                for (int m = 2; m < PM_P; ++m) t[m] = fma(x2, t[m - 1], -t[m - 2]);
            } else {
#pragma unroll
                for (int m = 0; m < PM_P; ++m) t[m] = 0.0;
            }
#pragma unroll
            for (int m = 0; m < PM_P; m += 2) *reinterpret_cast<double2*>(Ts + lane * PMK_LDT + m) = make_double2(t[m], t[m + 1]);
            __syncwarp();
            const int nk = (int)min((int64_t)8, (s1 - base + 3) / 4);       // k-steps of 4 spikes
            for (int j = 0; j < nk; ++j) {
                const double a0 = Ts[(4 * j + tg) * PMK_LDT + g], a1 = Ts[(4 * j + tg) * PMK_LDT + g + 8];
                const int bs = __shfl_sync(0xffffffffu, pb, 4 * j + tg);
#pragma unroll
                for (int nt = 0; nt < NT; ++nt) {
                    const double ind = (bs == 8 * nt + g) ? 1.0 : 0.0;      // B[k = spike tg][n = panel g]
                    pm_dmma(c[0][nt][0], c[0][nt][1], a0, ind);
                    pm_dmma(c[1][nt][0], c[1][nt][1], a1, ind);
                }
            }
            __syncwarp();                                 // the tile is rewritten by the next 32 spikes
        }
        // tau = D c: c from accumulator layout to B-fragment layout through shared memory
#pragma unroll
        for (int it = 0; it < 2; ++it)
#pragma unroll
            for (int nt = 0; nt < NT; ++nt)
                *reinterpret_cast<double2*>(Cs + (8 * it + g) * LDC + 8 * nt + 2 * tg) = make_double2(c[it][nt][0], c[it][nt][1]);
        __syncwarp();
#pragma unroll
        for (int nt = 0; nt < NT; ++nt) {
            double ta[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const double bv = Cs[(4 * ks + tg) * LDC + 8 * nt + g];
                pm_dmma(ta[0][0], ta[0][1], dfrag[0][ks], bv);
                pm_dmma(ta[1][0], ta[1][1], dfrag[1][ks], bv);
            }
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int pbn = 8 * nt + 2 * tg + e;      // panel; node 8 it + g
                if (pbn < dm.pm_B) {
                    out[pbn * PM_P + g] = ta[0][e];
                    out[pbn * PM_P + 8 + g] = ta[1][e];
                }
            }
        }
        __syncwarp();                                     // Cs is rewritten by the next segment
    }
}

// ------------------------------------------------------------------------------------------
// mun[r][k][i] = sum_j alpha_j kappa_k(t_i - z_j): CTA per (trial, latent), thread <-> node
// ------------------------------------------------------------------------------------------
constexpr int PMN_THREADS = 128;

__global__ void __launch_bounds__(PMN_THREADS) panel_nodal_means_kernel(svgpfa_dims dm, svgpfa_buffers bf) {
    __shared__ double zs[SVGPFA_MAX_M + 4], as[SVGPFA_MAX_M + 4];
    __shared__ double etab[64];
    __shared__ double2 sctab[SVGPFA_SC_ENTRIES];
    svgpfa_load_exp_tab64(etab);
    const int r = dm.r0 + blockIdx.x, k = blockIdx.y;
    const svgpfa_latent_desc ds = bf.desc[k];
    const int M = ds.M, NB = dm.pm_B * PM_P;
    const KConst kc = make_kconst(ds, bf.theta, bf.kscale, k);
    if (kc.type == SVGPFA_KERNEL_PERIODIC) svgpfa_load_sincos_tab<1>(sctab);
    for (int j = threadIdx.x; j < SVGPFA_MAX_M + 4; j += blockDim.x) {
        zs[j] = j < M ? bf.Z[(size_t)dm.R * ds.moff + (size_t)r * M + j] : 0.0;
        as[j] = j < M ? bf.alpha[(size_t)r * dm.KM + ds.moff + j] : 0.0;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < NB; i += blockDim.x) {
        const double t = pm_node_time(dm, i);
        double mu = 0.0;
#pragma unroll 1
        for (int j0 = 0; j0 < M; j0 += 4) {
            double dl[4], kv[4], qq[4], s2x[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) dl[e] = t - zs[j0 + e];
            kappa_vals_n<4>(kc, dl, etab, sctab, kv, qq, s2x);
#pragma unroll
            for (int e = 0; e < 4; ++e) mu = fma(kv[e], as[j0 + e], mu);          // as = 0 beyond M
        }
        bf.pm_mun[((size_t)r * dm.K + k) * NB + i] = mu;
    }
}

// ------------------------------------------------------------------------------------------
// mt[r][k][i] = sum_n C[n,k] tau[r][n][i]   (K x N)(N x NB) per trial on the FP64 tensor path.
// Persistent CTA, 8 warps; warp w owns the node columns [64 NCT ... ) / 8 * w (NCT 8-column tiles) for all K rows,
// so every tau element is read exactly once, straight from global memory (no reuse: no shared-memory staging);
// C^T is the A operand from shared memory (leading dimension = 8 or 24 mod 32: conflict-free fragments).
// ------------------------------------------------------------------------------------------
constexpr int PMG_THREADS = 256;

__host__ __device__ inline int pm_ldc(int KT) { const int kp = 8 * KT; return (kp % 32 == 8 || kp % 32 == 24) ? kp : kp + 8; }

template <int KT, int NCT>
__global__ void __launch_bounds__(PMG_THREADS, 2) panel_weights_kernel(svgpfa_dims dm, svgpfa_buffers bf) {
    extern __shared__ double pg_sm[];                     // C [N4][LDC], rows >= N and columns >= K zero
    constexpr int LDC = (8 * KT % 32 == 8 || 8 * KT % 32 == 24) ? 8 * KT : 8 * KT + 8;
    const int N = dm.N, K = dm.K, NB = dm.pm_B * PM_P, N4 = (N + 3) / 4 * 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tg = lane & 3;
    for (int idx = tid; idx < N4 * LDC; idx += PMG_THREADS) {
        const int n = idx / LDC, kk = idx - n * LDC;
        pg_sm[idx] = (n < N && kk < K) ? bf.C[(size_t)n * K + kk] : 0.0;
    }
    __syncthreads();
    const int nt = dm.rn ? dm.rn : dm.R;
    // Software pipeline over batches of PW k-steps (PW * NCT tau loads per lane): the loads of batch i + 1 -- of the
    // NEXT trial after a trial's last batch -- are issued before the mma of batch i.  Round-2 history: with the loads
    // left to the compiler long-scoreboard was the top stall; batched loads (issue all, wait, compute) left the kernel
    // alternating between a load phase and a compute phase (HBM 43 % + FP64 pipe 46 % of the time, summing to one).
    constexpr int PW = NCT <= 1 ? 16 : NCT == 2 ? 8 : NCT == 3 ? 5 : NCT == 4 ? 4 : 2;
    const int nks = N4 / 4, nbe = ((nks + PW - 1) / PW + 1) & ~1;          // batches per trial, padded to an even count
    double acc[KT][NCT][2];
    auto load = [&](double (&b)[PW][NCT], int rl, int ks0) {
        const double* tau = bf.pm_tau + (size_t)(dm.r0 + rl) * N * NB + 8 * NCT * warp + g;
#pragma unroll
        for (int u = 0; u < PW; ++u) {
            const int n = 4 * (ks0 + u) + tg;
#pragma unroll
            for (int c = 0; c < NCT; ++c) b[u][c] = n < N ? __ldg(tau + (size_t)n * NB + 8 * c) : 0.0;
        }
    };
    auto compute = [&](const double (&b)[PW][NCT], int ks0) {
#pragma unroll
        for (int u = 0; u < PW; ++u) {
            const int n = 4 * (ks0 + u) + tg;
            if (4 * (ks0 + u) >= N4) break;                // warp-uniform
            double a[KT];
#pragma unroll
            for (int kt = 0; kt < KT; ++kt) a[kt] = pg_sm[n * LDC + 8 * kt + g];
#pragma unroll
            for (int kt = 0; kt < KT; ++kt)
#pragma unroll
                for (int c = 0; c < NCT; ++c) pm_dmma(acc[kt][c][0], acc[kt][c][1], a[kt], b[u][c]);
        }
    };
    double b0[PW][NCT], b1[PW][NCT];
    int rl = blockIdx.x;
    if (rl < nt) load(b0, rl, 0);
    for (; rl < nt; rl += gridDim.x) {
        const int r = dm.r0 + rl;
#pragma unroll
        for (int a = 0; a < KT; ++a)
#pragma unroll
            for (int c = 0; c < NCT; ++c) acc[a][c][0] = acc[a][c][1] = 0.0;
        for (int bi = 0; bi < nbe; bi += 2) {
            load(b1, rl, (bi + 1) * PW);
            compute(b0, bi * PW);
            if (bi + 2 < nbe) load(b0, rl, (bi + 2) * PW);
            else if (rl + (int)gridDim.x < nt) load(b0, rl + gridDim.x, 0);
            compute(b1, (bi + 1) * PW);
        }
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) {
            const int kk = 8 * kt + g;
            if (kk < K) {
#pragma unroll
                for (int c = 0; c < NCT; ++c)
                    *reinterpret_cast<double2*>(bf.pm_mt + ((size_t)r * K + kk) * NB + 8 * (NCT * warp + c) + 2 * tg) =
                        make_double2(acc[kt][c][0], acc[kt][c][1]);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// dC[n][k] += sum_r sum_i tau[r][n][i] mun[r][k][i]   (N x R NB)(R NB x K): the reduction runs over trials and nodes.
// CTA = (block of 256 neurons, range of trials), 8 warps; warp w owns neuron rows 32 w ... 32 w + 31 of the block
// (four 8-row tiles) for all K columns and keeps its accumulators in registers over the whole trial range; tau is
// the A operand straight from global memory (each element read once), mun of the current trial the B operand from
// shared memory.  One atomic per (neuron, latent) and CTA at the end.
// ------------------------------------------------------------------------------------------
constexpr int PMD_ROWS = 256;

template <int KT>
__global__ void __launch_bounds__(PMG_THREADS, 2) panel_dC_kernel(svgpfa_dims dm, svgpfa_buffers bf, double* __restrict__ gC) {
    extern __shared__ double pd_sm[];                     // mun of one trial, [8 KT][NB + 4], rows >= K zero
    const int N = dm.N, K = dm.K, NB = dm.pm_B * PM_P, LDM = NB + 4;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tg = lane & 3;
    const int row0 = blockIdx.x * PMD_ROWS + 32 * warp;
    const int nt = dm.rn ? dm.rn : dm.R;
    double acc[4][KT][2];
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
        for (int kt = 0; kt < KT; ++kt) acc[t][kt][0] = acc[t][kt][1] = 0.0;
    for (int idx = tid; idx < 8 * KT * LDM; idx += PMG_THREADS) pd_sm[idx] = 0.0;      // padding rows / columns stay zero
    // Software pipeline over batches of PW k-steps (= 64 nodes of one 8-row tile; a trial is 4 NB / 64 batches, an even
    // count): the tau loads of the next batch -- of the NEXT trial after a trial's last batch, ahead of the barrier that
    // swaps mun -- are issued before the mma of the current one (see panel_weights_kernel).
    constexpr int PW = 16;
    const int bpt = NB / 64, nbatch = 4 * bpt;            // batches per row tile, per trial
    auto load = [&](double (&a)[PW], int rl, int j) {
        const int t = j / bpt, ks0 = (j - t * bpt) * PW;
        const int n = row0 + 8 * t + g;
        if (row0 + 8 * t >= N) return;                    // warp-uniform: rows past the matrix, never computed on
        const double* tau = bf.pm_tau + ((size_t)(dm.r0 + rl) * N + (n < N ? n : N - 1)) * NB + tg + 4 * ks0;
#pragma unroll
        for (int u = 0; u < PW; ++u) a[u] = n < N ? __ldg(tau + 4 * u) : 0.0;
    };
    auto compute = [&](const double (&a)[PW], int j) {
        const int t = j / bpt, ks0 = (j - t * bpt) * PW;
        if (row0 + 8 * t >= N) return;
#pragma unroll
        for (int tt = 0; tt < 4; ++tt) {                  // static accumulator index
            if (tt == t) {
#pragma unroll
                for (int u = 0; u < PW; ++u)
#pragma unroll
                    for (int kt = 0; kt < KT; ++kt)
                        pm_dmma(acc[tt][kt][0], acc[tt][kt][1], a[u], pd_sm[(8 * kt + g) * LDM + 4 * (ks0 + u) + tg]);
            }
        }
    };
    double a0[PW], a1[PW];
    int rl = blockIdx.y;
    if (rl < nt) load(a0, rl, 0);
    for (; rl < nt; rl += gridDim.y) {
        const int r = dm.r0 + rl;
        __syncthreads();                                  // the previous trial's mun has been consumed
        for (int idx = tid; idx < K * NB; idx += PMG_THREADS) {
            const int kk = idx / NB, i = idx - kk * NB;
            pd_sm[kk * LDM + i] = bf.pm_mun[((size_t)r * K + kk) * NB + i];
        }
        __syncthreads();
        for (int j = 0; j < nbatch; j += 2) {
            load(a1, rl, j + 1);
            compute(a0, j);
            if (j + 2 < nbatch) load(a0, rl, j + 2);
            else if (rl + (int)gridDim.y < nt) load(a0, rl + gridDim.y, 0);
            compute(a1, j + 1);
        }
    }
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int n = row0 + 8 * t + g;
        if (n < N) {
#pragma unroll
            for (int kt = 0; kt < KT; ++kt)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const int kk = 8 * kt + 2 * tg + e;
                    if (kk < K) atomicAdd(gC + (size_t)n * K + kk, acc[t][kt][e]);
                }
        }
    }
}

// ------------------------------------------------------------------------------------------
// abar_spk_j = sum_i mt_i kappa(t_i - z_j) and, with KGRAD, the z_j / theta adjoints: warp per (trial, latent),
// lane <-> inducing point, four nodes per iteration (interleaved evaluation chains, kappa_vals_n).
// ------------------------------------------------------------------------------------------
constexpr int PMA_WARPS = 4;

template <bool KGRAD>
__global__ void __launch_bounds__(32 * PMA_WARPS) panel_adjoint_kernel(svgpfa_dims dm, svgpfa_buffers bf, uint32_t flags, int nprob) {
    extern __shared__ double pa_sm[];                     // node times [NB] | per warp: node weights [NB]
    __shared__ double etab[64];
    __shared__ double2 sctab[SVGPFA_SC_ENTRIES];
    const int NB = dm.pm_B * PM_P;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    svgpfa_load_exp_tab64(etab);
    svgpfa_load_sincos_tab<1>(sctab);
    double* tn = pa_sm;
    double* mw = pa_sm + NB + (size_t)warp * NB;
    for (int i = threadIdx.x; i < NB; i += blockDim.x) tn[i] = pm_node_time(dm, i);
    __syncthreads();
    const int prob = blockIdx.x * PMA_WARPS + warp;
    if (prob >= nprob) return;                            // whole warps leave; no block barrier below
    const int rl = prob / dm.K, k = prob - rl * dm.K, r = dm.r0 + rl;
    const svgpfa_latent_desc ds = bf.desc[k];
    const int M = ds.M;
    const KConst kc = make_kconst(ds, bf.theta, bf.kscale, k);
    const double* mt = bf.pm_mt + ((size_t)r * dm.K + k) * NB;
    for (int i = lane; i < NB; i += 32) mw[i] = mt[i];
    __syncwarp();
    const size_t vo = (size_t)r * dm.KM + ds.moff;
    double th0 = 0.0, th1 = 0.0;
    for (int jb = 0; jb < M; jb += 32) {
        const int j = jb + lane;
        const bool in = j < M;
        const double zj = in ? bf.Z[(size_t)dm.R * ds.moff + (size_t)r * M + j] : 0.0;
        const double aj = in ? bf.alpha[vo + j] : 0.0;
        double abar = 0.0, dz_raw = 0.0, t0_raw = 0.0, t1_raw = 0.0;
#pragma unroll 1
        for (int i0 = 0; i0 < NB; i0 += 4) {
            double dl[4], kv[4], qq[4], s2x[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) dl[e] = tn[i0 + e] - zj;
            kappa_vals_n<4>(kc, dl, etab, sctab, kv, qq, s2x);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const double gk = mw[i0 + e] * kv[e];
                abar += gk;
                if (KGRAD) {
                    if (kc.type == SVGPFA_KERNEL_EXPQUAD) {
                        dz_raw = fma(gk, dl[e], dz_raw);
                        t0_raw = fma(gk, qq[e], t0_raw);
                    } else {
                        const double gs = gk * s2x[e];
                        dz_raw += gs;
                        t0_raw = fma(gk, qq[e], t0_raw);
                        t1_raw = fma(gs, dl[e], t1_raw);
                    }
                }
            }
        }
        if (in) {
            bf.abar_spk[vo + j] = abar;                   // sole writer on this path (the direct kernel accumulates)
            // d delta / d z = -1; dkappa/ddelta = kappa (delta | sin 2x) dd (kappa_grad, common.cuh)
            if (KGRAD && (flags & SVGPFA_GRAD_INDLOCS)) atomicAdd(bf.dz_acc + vo + j, -aj * kc.dd * dz_raw);
            th0 = fma(aj * kc.dl, t0_raw, th0);
            th1 = fma(aj * kc.dp, t1_raw, th1);
        }
    }
    if (KGRAD && (flags & SVGPFA_GRAD_KERNEL)) {
        th0 = warp_sum(th0);
        th1 = warp_sum(th1);
        if (lane == 0) {
            double* dth = bf.dth_part + (size_t)r * dm.TH + ds.thoff;
            atomicAdd(dth, th0);                          // zeroed by the caller; the quadrature adjoint adds too
            if (ds.nth > 1) atomicAdd(dth + 1, th1);
        }
    }
}

template <int KT>
int launch_weights(const svgpfa_dims* dims, const svgpfa_buffers* buf, cudaStream_t st) {
    const int N4 = (dims->N + 3) / 4 * 4;
    const size_t smem = sizeof(double) * (size_t)N4 * pm_ldc(KT);
    if (smem > 200 * 1024) return svgpfa_set_error(SVGPFA_E_UNSUPPORTED, "spike_panel: N too large for the C tile", cudaSuccess);
    const int nt = svgpfa_ntrials(dims);
    int grid = 2 * svgpfa_sm_count();
    if (smem > 110 * 1024) grid = svgpfa_sm_count();
    if (grid > nt) grid = nt;
#define PM_LAUNCH_W(NCT)                                                                  \
    do {                                                                                  \
        SVGPFA_ENSURE_SMEM(smem, panel_weights_kernel<KT, NCT>);                          \
        panel_weights_kernel<KT, NCT><<<grid, PMG_THREADS, smem, st>>>(*dims, *buf);      \
    } while (0)
    switch (dims->pm_B) {                 // NB = 64 NCT node columns, 8 NCT per warp
        case 4: PM_LAUNCH_W(1); break;
        case 8: PM_LAUNCH_W(2); break;
        case 12: PM_LAUNCH_W(3); break;
        case 16: PM_LAUNCH_W(4); break;
        case 24: PM_LAUNCH_W(6); break;
        default: PM_LAUNCH_W(8); break;
    }
#undef PM_LAUNCH_W
    return SVGPFA_OK;
}

template <int KT>
int launch_dC(const svgpfa_dims* dims, const svgpfa_buffers* buf, double* out, cudaStream_t st) {
    const int NB = dims->pm_B * PM_P;
    const size_t smem = sizeof(double) * (size_t)8 * KT * (NB + 4);
    const int nblk = (dims->N + PMD_ROWS - 1) / PMD_ROWS, nt = svgpfa_ntrials(dims);
    int gy = (2 * svgpfa_sm_count() + nblk - 1) / nblk;      // one wave of two resident CTAs per SM (96 registers)
    if (gy > nt) gy = nt;
    SVGPFA_ENSURE_SMEM(smem, panel_dC_kernel<KT>);
    panel_dC_kernel<KT><<<dim3(nblk, gy), PMG_THREADS, smem, st>>>(*dims, *buf, out);
    return SVGPFA_OK;
}

int launch_dC_any(const svgpfa_dims* dims, const svgpfa_buffers* buf, double* out, cudaStream_t st) {
    switch ((dims->K + 7) / 8) {
        case 1: return launch_dC<1>(dims, buf, out, st);
        case 2: return launch_dC<2>(dims, buf, out, st);
        case 3: return launch_dC<3>(dims, buf, out, st);
        case 4: return launch_dC<4>(dims, buf, out, st);
        default: return launch_dC<5>(dims, buf, out, st);
    }
}

bool panel_args_ok(const svgpfa_dims* d, const svgpfa_buffers* b) {
    const int B = d ? d->pm_B : 0;
    return d && b && (B == 4 || B == 8 || B == 12 || B == 16 || B == 24 || B == 32) && d->pm_w > 0.0 && b->pm_tau && d->K <= 40;
}

}  // namespace

extern "C" int svgpfa_panel_moments(const svgpfa_dims* dims, const svgpfa_buffers* buf, void* stream) {
    if (!panel_args_ok(dims, buf)) return svgpfa_set_error(SVGPFA_E_ARG, "panel_moments", cudaSuccess);
    const int nt = svgpfa_ntrials(dims);
    if (nt == 0 || dims->N == 0) return SVGPFA_OK;
    long n_chunks = ((long)svgpfa_sm_count() * 16 + nt - 1) / nt;      // enough CTAs to fill the machine when R is small
    if (n_chunks < 1) n_chunks = 1;
    if (n_chunks > (dims->N + PMK_WARPS - 1) / PMK_WARPS) n_chunks = (dims->N + PMK_WARPS - 1) / PMK_WARPS;
    const int chunk = (int)((dims->N + n_chunks - 1) / n_chunks);
    n_chunks = (dims->N + chunk - 1) / chunk;
    const unsigned grid = (unsigned)(nt * n_chunks);
    cudaStream_t st = (cudaStream_t)stream;
    switch ((dims->pm_B + 7) / 8) {
        case 1: panel_moments_kernel<1><<<grid, 32 * PMK_WARPS, 0, st>>>(*dims, *buf, (int)n_chunks, chunk); break;
        case 2: panel_moments_kernel<2><<<grid, 32 * PMK_WARPS, 0, st>>>(*dims, *buf, (int)n_chunks, chunk); break;
        case 3: panel_moments_kernel<3><<<grid, 32 * PMK_WARPS, 0, st>>>(*dims, *buf, (int)n_chunks, chunk); break;
        default: panel_moments_kernel<4><<<grid, 32 * PMK_WARPS, 0, st>>>(*dims, *buf, (int)n_chunks, chunk); break;
    }
    SVGPFA_CHECK_LAUNCH("panel_moments");
    return SVGPFA_OK;
}

// gsum[n][k] = sum over the shard's spikes of neuron n of the latent mean mu_k(t_s), through the panel moments: what the
// cached-statistics path needs from the spike times (see spike_gather_kernel in spike.cu) without the S x K array.
extern "C" int svgpfa_panel_neuron_sums(const svgpfa_dims* dims, const svgpfa_buffers* buf, void* stream) {
    if (!panel_args_ok(dims, buf) || !buf->pm_mun || !buf->gsum)
        return svgpfa_set_error(SVGPFA_E_ARG, "panel_neuron_sums", cudaSuccess);
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemsetAsync(buf->gsum, 0, sizeof(double) * (size_t)dims->N * dims->K, st);
    const int nt = svgpfa_ntrials(dims);
    if (nt == 0 || dims->N == 0) return SVGPFA_OK;
    panel_nodal_means_kernel<<<dim3(nt, dims->K), PMN_THREADS, 0, st>>>(*dims, *buf);
    SVGPFA_CHECK_LAUNCH("panel_nodal_means");
    const int rc = launch_dC_any(dims, buf, buf->gsum, st);
    if (rc) return rc;
    SVGPFA_CHECK_LAUNCH("panel_dC (neuron sums)");
    return SVGPFA_OK;
}

extern "C" int svgpfa_spike_panel_fwd_bwd(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, void* stream) {
    if (!panel_args_ok(dims, buf) || !buf->pm_mt || !buf->pm_mun)
        return svgpfa_set_error(SVGPFA_E_ARG, "spike_panel_fwd_bwd", cudaSuccess);
    const int nt = svgpfa_ntrials(dims);
    if (nt == 0 || dims->N == 0) return SVGPFA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (flags & SVGPFA_REBUILD_PANELS) { rc = svgpfa_panel_moments(dims, buf, stream); if (rc) return rc; }
    const bool need_emb = flags & SVGPFA_GRAD_EMBEDDING;
    const bool kgrad = flags & (SVGPFA_GRAD_KERNEL | SVGPFA_GRAD_INDLOCS);
    const int NB = dims->pm_B * PM_P, KT = (dims->K + 7) / 8;
    if (need_emb) {
        panel_nodal_means_kernel<<<dim3(nt, dims->K), PMN_THREADS, 0, st>>>(*dims, *buf);
        SVGPFA_CHECK_LAUNCH("panel_nodal_means");
    }
    switch (KT) {
        case 1: rc = launch_weights<1>(dims, buf, st); break;
        case 2: rc = launch_weights<2>(dims, buf, st); break;
        case 3: rc = launch_weights<3>(dims, buf, st); break;
        case 4: rc = launch_weights<4>(dims, buf, st); break;
        default: rc = launch_weights<5>(dims, buf, st); break;
    }
    if (rc) return rc;
    SVGPFA_CHECK_LAUNCH("panel_weights");
    if (need_emb) {
        rc = launch_dC_any(dims, buf, buf->shared + SVGPFA_SHARED_HDR, st);
        if (rc) return rc;
        SVGPFA_CHECK_LAUNCH("panel_dC");
    }
    const int nprob = nt * dims->K;
    const size_t smem = sizeof(double) * (size_t)NB * (1 + PMA_WARPS);
    const int blocks = (nprob + PMA_WARPS - 1) / PMA_WARPS;
    if (kgrad) panel_adjoint_kernel<true><<<blocks, 32 * PMA_WARPS, smem, st>>>(*dims, *buf, flags, nprob);
    else panel_adjoint_kernel<false><<<blocks, 32 * PMA_WARPS, smem, st>>>(*dims, *buf, flags, nprob);
    SVGPFA_CHECK_LAUNCH("panel_adjoint");
    return SVGPFA_OK;
}

// Legendre-quadrature part of the expected log-likelihood (sm_100a, float64):
//   quad_latent_fwd_kernel   mu, var of every latent at every quadrature point        (R,Q,K)
//   quad_embed_kernel        h = C x + d, exp(mean + var/2), weighted integral, dC, dd, mubar, varbar
//   quad_latent_bwd_kernel   adjoints of the latent posterior: A_q, abar_q, dz_acc, dth_part
// Ktz (R,Q,M), Kzz^-1 Kzt (R,M,Q) and eLinkValues (R,Q,N) of the reference are never materialised
// (stats/kernelsMatricesStore.py:186-195, stats/svPosteriorOnLatents.py:185-216,
//  stats/svEmbedding.py:80-84, stats/expectedLogLikelihood.py:107-135,205-208).
#include "common.cuh"

namespace {

// ======================================================================================
// latent posterior at quadrature points: thread <-> quadrature point, one CTA per (trial, latent)
// ======================================================================================
constexpr int QL_TQ = 128;            // quadrature points per pass (= threads per CTA)
constexpr int QL_TQS = QL_TQ + 1;     // odd row stride of the per-point vectors

struct QLSmem {
    double *LiT, *Li, *X, *XT;        // MP x MP each (zero padded), row stride MP
    double *ks, *vs, *us;             // MP x QL_TQS
    double *al, *zs;                  // MP
    double *mb, *vb;                  // QL_TQ
};

__host__ __device__ inline size_t ql_smem_bytes(int MP, bool bwd) {
    size_t n = (size_t)(bwd ? 4 : 2) * MP * MP + (size_t)(bwd ? 3 : 2) * MP * QL_TQS + 2 * MP + 2 * QL_TQ;
    return n * sizeof(double);
}

__device__ __forceinline__ QLSmem ql_carve(double* sm, int MP, bool bwd) {
    QLSmem s;
    s.LiT = sm;
    s.X = s.LiT + MP * MP;
    double* p = s.X + MP * MP;
    if (bwd) { s.Li = p; s.XT = p + MP * MP; p += 2 * MP * MP; } else { s.Li = nullptr; s.XT = nullptr; }
    s.ks = p; p += MP * QL_TQS;
    s.vs = p; p += MP * QL_TQS;
    if (bwd) { s.us = p; p += MP * QL_TQS; } else s.us = nullptr;
    s.al = p; p += MP;
    s.zs = p; p += MP;
    s.mb = p; p += QL_TQ;
    s.vb = p;
    return s;
}

// out_i = sum_j Mt[j][i] in_j, i.e. out = Mat * in with Mat given TRANSPOSED (Mt[j*MP+i] = Mat[i][j]).
// lower == true : Mat lower-triangular  (j <= i)     -> j in [0, i0+3]
// lower == false: Mat upper-triangular  (j >= i)     -> j in [i0, M)
// 4 outputs per pass; Mt rows are read as 2 x 16-byte broadcasts, `in` is a per-thread column.
template <bool LOWER, class Out>
__device__ __forceinline__ void tri_matvec4(const double* __restrict__ Mt, const double* __restrict__ in, int MP,
                                            int M, int tq, Out out) {
    for (int i0 = 0; i0 < MP; i0 += 4) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        const int jb = LOWER ? 0 : i0;
        const int je = LOWER ? min(i0 + 4, M) : M;
        for (int j = jb; j < je; ++j) {
            const double x = in[j * QL_TQS + tq];
            const double2 m01 = *reinterpret_cast<const double2*>(Mt + j * MP + i0);
            const double2 m23 = *reinterpret_cast<const double2*>(Mt + j * MP + i0 + 2);
            a0 = fma(m01.x, x, a0);
            a1 = fma(m01.y, x, a1);
            a2 = fma(m23.x, x, a2);
            a3 = fma(m23.y, x, a3);
        }
        out(i0, a0, a1, a2, a3);
    }
}

__device__ __forceinline__ void ql_load_mats(const QLSmem& s, const svgpfa_dims& dm, const svgpfa_buffers& bf,
                                             const svgpfa_latent_desc& ds, int r, int MP, bool bwd) {
    const int M = ds.M;
    const size_t mo = (size_t)r * dm.MM + ds.mmoff;
    for (int idx = threadIdx.x; idx < MP * MP; idx += blockDim.x) {
        const int i = idx / MP, j = idx - i * MP;
        const bool in = (i < M) && (j < M);
        const double li = in ? bf.Li[mo + (size_t)i * M + j] : 0.0;
        const double x = in ? bf.X[mo + (size_t)i * M + j] : 0.0;
        s.LiT[j * MP + i] = li;
        s.X[i * MP + j] = x;
        if (bwd) { s.Li[i * MP + j] = li; s.XT[j * MP + i] = x; }
    }
    const double* zg = bf.Z + (size_t)dm.R * ds.moff + (size_t)r * M;
    const size_t vo = (size_t)r * dm.KM + ds.moff;
    for (int i = threadIdx.x; i < MP; i += blockDim.x) {
        s.zs[i] = (i < M) ? zg[i] : 0.0;
        s.al[i] = (i < M) ? bf.alpha[vo + i] : 0.0;
    }
}

__global__ void __launch_bounds__(QL_TQ) quad_latent_fwd_kernel(svgpfa_dims dm, svgpfa_buffers bf) {
    extern __shared__ __align__(16) double sm[];
    const int r = blockIdx.x, k = blockIdx.y;
    const svgpfa_latent_desc ds = bf.desc[k];
    const int M = ds.M, MP = round_up(M, 4);
    const QLSmem s = ql_carve(sm, MP, false);
    const KConst kc = make_kconst(ds, bf.theta, bf.kscale, k);
    ql_load_mats(s, dm, bf, ds, r, MP, false);
    __syncthreads();
    const int tq = threadIdx.x;
    for (int q0 = 0; q0 < dm.Q; q0 += QL_TQ) {
        const int q = q0 + tq;
        if (q < dm.Q) {
            const double t = bf.tq[(size_t)r * dm.Q + q];
            double mu = 0.0;
            for (int j = 0; j < M; ++j) {
                const double kv = kappa_val(kc, t - s.zs[j]);
                s.ks[j * QL_TQS + tq] = kv;
                mu = fma(kv, s.al[j], mu);
            }
            double vv = 0.0, uu = 0.0;
            // v = Li k
            tri_matvec4<true>(s.LiT, s.ks, MP, M, tq, [&](int i0, double a0, double a1, double a2, double a3) {
                s.vs[(i0 + 0) * QL_TQS + tq] = a0;
                s.vs[(i0 + 1) * QL_TQS + tq] = a1;
                s.vs[(i0 + 2) * QL_TQS + tq] = a2;
                s.vs[(i0 + 3) * QL_TQS + tq] = a3;
                vv += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
            });
            // u = X^T v  (X^T upper-triangular; its transpose is X, row-major)
            tri_matvec4<false>(s.X, s.vs, MP, M, tq, [&](int, double a0, double a1, double a2, double a3) {
                uu += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
            });
            const size_t o = ((size_t)r * dm.Q + q) * dm.K + k;
            bf.mu_q[o] = mu;
            bf.var_q[o] = kc.s2 - vv + uu;
        }
    }
}

__global__ void __launch_bounds__(QL_TQ) quad_latent_bwd_kernel(svgpfa_dims dm, svgpfa_buffers bf, uint32_t flags) {
    extern __shared__ __align__(16) double sm[];
    __shared__ double red[32];
    const int r = blockIdx.x, k = blockIdx.y;
    const svgpfa_latent_desc ds = bf.desc[k];
    const int M = ds.M, MP = round_up(M, 4);
    const bool need_kz = flags & (SVGPFA_GRAD_KERNEL | SVGPFA_GRAD_INDLOCS);
    const QLSmem s = ql_carve(sm, MP, true);
    const KConst kc = make_kconst(ds, bf.theta, bf.kscale, k);
    ql_load_mats(s, dm, bf, ds, r, MP, true);
    const int tq = threadIdx.x;
    // phase-2 ownership: 4x4 tiles of the lower triangle of A, G groups splitting the q range
    const int nt = MP / 4, ntile = nt * (nt + 1) / 2;
    const int G = max(1, (int)blockDim.x / ntile);
    const int my_tile = tq % ntile, my_g = tq / ntile;
    const bool has_tile = (tq < ntile * G);
    int ti = 0, tj = 0;
    {   // decode my_tile -> (ti >= tj)
        int t = my_tile, row = 0;
        while (t >= row + 1) { t -= row + 1; ++row; }
        ti = row; tj = t;
    }
    // M > 44: more tiles than threads (G == 1); threads tq < ntile - blockDim own a second tile
    const int my_tile2 = tq + (int)blockDim.x;
    const bool has_tile2 = (G == 1) && (my_tile2 < ntile);
    int ti2 = 0, tj2 = 0;
    if (has_tile2) {
        int t = my_tile2, row = 0;
        while (t >= row + 1) { t -= row + 1; ++row; }
        ti2 = row; tj2 = t;
    }
    double acc[16], acc2[16];
#pragma unroll
    for (int e = 0; e < 16; ++e) { acc[e] = 0.0; acc2[e] = 0.0; }
    double ab_acc = 0.0, dz_accum = 0.0, th0 = 0.0, th1 = 0.0;
    const size_t part_stride = (size_t)dm.R * dm.Q * dm.K;
    __syncthreads();
    for (int q0 = 0; q0 < dm.Q; q0 += QL_TQ) {
        const int q = q0 + tq;
        const bool valid = q < dm.Q;
        double t = 0.0, mbar = 0.0, vbar = 0.0;
        if (valid) {
            t = bf.tq[(size_t)r * dm.Q + q];
            const size_t o = ((size_t)r * dm.Q + q) * dm.K + k;
            for (int p = 0; p < dm.n_ntiles; ++p) {
                mbar += bf.mubar_part[p * part_stride + o];
                vbar += bf.varbar_part[p * part_stride + o];
            }
        }
        s.mb[tq] = mbar;
        s.vb[tq] = vbar;
        for (int j = 0; j < MP; ++j) {
            const double kv = (valid && j < M) ? kappa_val(kc, t - s.zs[j]) : 0.0;
            s.ks[j * QL_TQS + tq] = kv;
        }
        tri_matvec4<true>(s.LiT, s.ks, MP, M, tq, [&](int i0, double a0, double a1, double a2, double a3) {
            s.vs[(i0 + 0) * QL_TQS + tq] = a0;
            s.vs[(i0 + 1) * QL_TQS + tq] = a1;
            s.vs[(i0 + 2) * QL_TQS + tq] = a2;
            s.vs[(i0 + 3) * QL_TQS + tq] = a3;
        });
        if (need_kz) {
            // u = X^T v ; w = X u - v ; kbar = 2 vbar Li^T w + mubar alpha ; g_j = kbar_j dkappa/ddelta
            tri_matvec4<false>(s.X, s.vs, MP, M, tq, [&](int i0, double a0, double a1, double a2, double a3) {
                s.us[(i0 + 0) * QL_TQS + tq] = a0;
                s.us[(i0 + 1) * QL_TQS + tq] = a1;
                s.us[(i0 + 2) * QL_TQS + tq] = a2;
                s.us[(i0 + 3) * QL_TQS + tq] = a3;
            });
            // w overwrites us only after the whole matvec (reads us) is done: stage in registers per 4-block
            // is not possible (w_i needs all u_j, j<=i), so write w into ks' slot? ks is still needed for
            // abar.  Use a two-step: w -> vs2 := us (safe: block i0 reads u_j for j <= i0+3 only, and
            // blocks are processed in DESCENDING order so that u_j, j <= i0+3, are still intact).
            for (int i0 = MP - 4; i0 >= 0; i0 -= 4) {
                double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
                const int je = min(i0 + 4, M);
                for (int j = 0; j < je; ++j) {
                    const double x = s.us[j * QL_TQS + tq];
                    const double2 m01 = *reinterpret_cast<const double2*>(s.XT + j * MP + i0);
                    const double2 m23 = *reinterpret_cast<const double2*>(s.XT + j * MP + i0 + 2);
                    a0 = fma(m01.x, x, a0);
                    a1 = fma(m01.y, x, a1);
                    a2 = fma(m23.x, x, a2);
                    a3 = fma(m23.y, x, a3);
                }
                s.us[(i0 + 0) * QL_TQS + tq] = a0 - s.vs[(i0 + 0) * QL_TQS + tq];
                s.us[(i0 + 1) * QL_TQS + tq] = a1 - s.vs[(i0 + 1) * QL_TQS + tq];
                s.us[(i0 + 2) * QL_TQS + tq] = a2 - s.vs[(i0 + 2) * QL_TQS + tq];
                s.us[(i0 + 3) * QL_TQS + tq] = a3 - s.vs[(i0 + 3) * QL_TQS + tq];
            }
            // kv = Li^T w (Li^T upper; its transpose is Li row-major); ascending blocks read w_i, i >= j0,
            // and write slot j0..j0+3 AFTER reading -> in place is safe.
            tri_matvec4<false>(s.Li, s.us, MP, M, tq, [&](int j0, double a0, double a1, double a2, double a3) {
                const double av[4] = {a0, a1, a2, a3};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = j0 + e;
                    double g = 0.0;
                    if (valid && j < M) {
                        const double kbar = 2.0 * vbar * av[e] + mbar * s.al[j];
                        double kv, dkd, d0, d1;
                        kappa_grad(kc, t - s.zs[j], kv, dkd, d0, d1);
                        g = -kbar * dkd;           // d delta / d z = -1
                        th0 = fma(kbar, d0, th0);
                        th1 = fma(kbar, d1, th1);
                    }
                    s.us[j * QL_TQS + tq] = g;
                }
            });
        }
        __syncthreads();
        // ---- phase 2: reductions over the quadrature points of this pass
        if (has_tile) {
            for (int qq = my_g; qq < QL_TQ; qq += G) {
                const double sv = s.vb[qq];
                double vi[4], vj[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    vi[e] = s.vs[(4 * ti + e) * QL_TQS + qq] * sv;
                    vj[e] = s.vs[(4 * tj + e) * QL_TQS + qq];
                }
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a * 4 + b] = fma(vi[a], vj[b], acc[a * 4 + b]);
            }
        }
        if (has_tile2) {
            for (int qq = 0; qq < QL_TQ; ++qq) {
                const double sv = s.vb[qq];
                double vi[4], vj[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    vi[e] = s.vs[(4 * ti2 + e) * QL_TQS + qq] * sv;
                    vj[e] = s.vs[(4 * tj2 + e) * QL_TQS + qq];
                }
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc2[a * 4 + b] = fma(vi[a], vj[b], acc2[a * 4 + b]);
            }
        }
        if (tq < M) {
            double sa = 0.0, sz = 0.0;
            for (int qq = 0; qq < QL_TQ; ++qq) {
                sa = fma(s.mb[qq], s.ks[tq * QL_TQS + qq], sa);
                if (need_kz) sz += s.us[tq * QL_TQS + qq];
            }
            ab_acc += sa;
            dz_accum += sz;
        }
        __syncthreads();
    }
    // ---- write-out: combine the G partial copies of every A tile through shared memory
    double* scratch = s.ks;               // >= ntile*16 doubles?  MP*QL_TQS >= (MP/4)(MP/4+1)/2*16 for MP >= 4
    const size_t mo = (size_t)r * dm.MM + ds.mmoff;
    for (int g = 0; g < G; ++g) {
        if (has_tile && my_g == g) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                if (g == 0) scratch[my_tile * 16 + e] = acc[e];
                else scratch[my_tile * 16 + e] += acc[e];
            }
        }
        __syncthreads();
    }
    if (has_tile2) {
#pragma unroll
        for (int e = 0; e < 16; ++e) scratch[my_tile2 * 16 + e] = acc2[e];
    }
    __syncthreads();
    for (int idx = tq; idx < ntile * 16; idx += blockDim.x) {
        const int tl = idx / 16, e = idx - tl * 16;
        int t2 = tl, row = 0;
        while (t2 >= row + 1) { t2 -= row + 1; ++row; }
        const int i = 4 * row + e / 4, j = 4 * t2 + (e & 3);
        if (i < M && j <= i) bf.A_q[mo + (size_t)i * M + j] = scratch[idx];
    }
    const size_t vo = (size_t)r * dm.KM + ds.moff;
    if (tq < M) {
        bf.abar_q[vo + tq] = ab_acc;
        if (need_kz) bf.dz_acc[vo + tq] = dz_accum;      // first writer of dz_acc (spike kernel adds later)
    }
    if (need_kz && (flags & SVGPFA_GRAD_KERNEL)) {
        const double s0 = block_sum(th0, red);
        const double s1 = block_sum(th1, red);
        if (tq == 0) {
            double* dth = bf.dth_part + (size_t)r * dm.TH + ds.thoff;
            dth[0] = s0;                                 // first writer of dth_part
            if (ds.nth > 1) dth[1] = s1;
        }
    }
}

// ======================================================================================
// embedding + exp link + integral, one CTA per (neuron tile, worker); workers stride over
// (trial, 16-point) items
// ======================================================================================
constexpr int EM_TN = SVGPFA_EMBED_TN;     // 128 neurons per tile
constexpr int EM_TNS = EM_TN + 2;          // row stride (even: 16-byte aligned rows)
constexpr int EM_TQ = 16;                  // quadrature points per item
constexpr int EM_THREADS = 256;

__host__ __device__ inline size_t em_smem_bytes(int K) {
    // CT[K][TNS], dC[K][TNS], G[TQ][TNS], muT[K][TQ], varT[K][TQ], w[TQ], dvec[TN]
    return sizeof(double) * ((size_t)2 * K * EM_TNS + (size_t)EM_TQ * EM_TNS + (size_t)2 * K * EM_TQ + EM_TQ + EM_TN);
}

__global__ void __launch_bounds__(EM_THREADS) quad_embed_kernel(svgpfa_dims dm, svgpfa_buffers bf, uint32_t flags) {
    extern __shared__ __align__(16) double sm[];
    __shared__ double red[32];
    const int K = dm.K, N = dm.N, Q = dm.Q;
    double* CT = sm;                               // [k][n]
    double* dCs = CT + (size_t)K * EM_TNS;         // [k][n]
    double* Gs = dCs + (size_t)K * EM_TNS;         // [q][n]
    double* muT = Gs + (size_t)EM_TQ * EM_TNS;     // [k][q]
    double* varT = muT + (size_t)K * EM_TQ;        // [k][q]
    double* ws = varT + (size_t)K * EM_TQ;         // [q]
    double* dvec = ws + EM_TQ;                     // [n]
    const int tid = threadIdx.x;
    const int tile = blockIdx.x, n0 = tile * EM_TN;
    const int nloc = tid & (EM_TN - 1), half = tid >> 7;       // 2 halves of 128 threads
    const int n = n0 + nloc;
    const bool need_emb = flags & SVGPFA_GRAD_EMBEDDING;
    const bool need_lat = flags & (SVGPFA_GRAD_POSTERIOR | SVGPFA_GRAD_KERNEL | SVGPFA_GRAD_INDLOCS);
    for (int idx = tid; idx < K * EM_TN; idx += EM_THREADS) {
        const int kk = idx / EM_TN, nn = idx - kk * EM_TN;
        CT[kk * EM_TNS + nn] = (n0 + nn < N) ? bf.C[(size_t)(n0 + nn) * K + kk] : 0.0;
        dCs[kk * EM_TNS + nn] = 0.0;
    }
    if (tid < EM_TN) dvec[tid] = (n0 + tid < N) ? bf.d[n0 + tid] : 0.0;
    double t1 = 0.0, dd_acc = 0.0;
    const int qtiles = (Q + EM_TQ - 1) / EM_TQ;
    const int nitems = dm.R * qtiles;
    const size_t part_off = (size_t)tile * dm.R * Q * K;
    __syncthreads();
    for (int it = blockIdx.y; it < nitems; it += gridDim.y) {
        const int r = it / qtiles, q0 = (it - r * qtiles) * EM_TQ;
        // stage mu, var (transposed) and weights of the 16 points
        for (int idx = tid; idx < EM_TQ * K; idx += EM_THREADS) {
            const int qq = idx / K, kk = idx - qq * K;
            const bool v = (q0 + qq) < Q;
            const size_t o = ((size_t)r * Q + q0 + qq) * K + kk;
            muT[kk * EM_TQ + qq] = v ? bf.mu_q[o] : 0.0;
            varT[kk * EM_TQ + qq] = v ? bf.var_q[o] : 0.0;
        }
        if (tid < EM_TQ) ws[tid] = (q0 + tid < Q) ? bf.wq[(size_t)r * Q + q0 + tid] : 0.0;
        __syncthreads();
        // ---- phase A: thread = (neuron, 8 points)
        {
            double h[8], sg[8];
            const double dn = dvec[nloc];
#pragma unroll
            for (int e = 0; e < 8; ++e) { h[e] = dn; sg[e] = 0.0; }
            const int qb = half * 8;
            for (int kk = 0; kk < K; ++kk) {
                const double c = CT[kk * EM_TNS + nloc], c2 = c * c;
#pragma unroll
                for (int e = 0; e < 8; e += 2) {
                    const double2 m2 = *reinterpret_cast<const double2*>(muT + kk * EM_TQ + qb + e);
                    const double2 v2 = *reinterpret_cast<const double2*>(varT + kk * EM_TQ + qb + e);
                    h[e] = fma(c, m2.x, h[e]);
                    h[e + 1] = fma(c, m2.y, h[e + 1]);
                    sg[e] = fma(c2, v2.x, sg[e]);
                    sg[e + 1] = fma(c2, v2.y, sg[e + 1]);
                }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const double w = (n < N) ? ws[qb + e] : 0.0;
                const double ev = w * exp(fma(0.5, sg[e], h[e]));
                t1 += ev;
                Gs[(qb + e) * EM_TNS + nloc] = -ev;
            }
        }
        __syncthreads();
        // ---- phase B: mubar[q][k] = sum_n G[q][n] C[n][k], varbar = 0.5 sum_n G C^2 (this tile's partial)
        if (need_lat) {
            for (int o = tid; o < EM_TQ * K; o += EM_THREADS) {
                const int qq = o / K, kk = o - qq * K;
                double sm_ = 0.0, sv_ = 0.0;
                const double* g = Gs + qq * EM_TNS;
                const double* c = CT + kk * EM_TNS;
#pragma unroll 4
                for (int nn = 0; nn < EM_TN; nn += 2) {
                    const double2 g2 = *reinterpret_cast<const double2*>(g + nn);
                    const double2 c2 = *reinterpret_cast<const double2*>(c + nn);
                    const double a = g2.x * c2.x, b = g2.y * c2.y;
                    sm_ += a + b;
                    sv_ = fma(a, c2.x, sv_);
                    sv_ = fma(b, c2.y, sv_);
                }
                if (q0 + qq < Q) {
                    const size_t oo = part_off + ((size_t)r * Q + q0 + qq) * K + kk;
                    bf.mubar_part[oo] = sm_;
                    bf.varbar_part[oo] = 0.5 * sv_;
                }
            }
        }
        // ---- phase C: dC[n][k] += sum_q G[q][n] (mu[q][k] + C[n][k] var[q][k]);  dd[n] += sum_q G[q][n]
        if (need_emb) {
            double g[EM_TQ];
            double gs = 0.0;
#pragma unroll
            for (int e = 0; e < EM_TQ; ++e) { g[e] = Gs[e * EM_TNS + nloc]; gs += g[e]; }
            if (half == 0) dd_acc += gs;
            for (int kk = half; kk < K; kk += 2) {
                const double c = CT[kk * EM_TNS + nloc];
                double a = 0.0;
#pragma unroll
                for (int e = 0; e < EM_TQ; e += 2) {
                    const double2 m2 = *reinterpret_cast<const double2*>(muT + kk * EM_TQ + e);
                    const double2 v2 = *reinterpret_cast<const double2*>(varT + kk * EM_TQ + e);
                    a = fma(g[e], fma(c, v2.x, m2.x), a);
                    a = fma(g[e + 1], fma(c, v2.y, m2.y), a);
                }
                dCs[kk * EM_TNS + nloc] += a;
            }
        }
        __syncthreads();
    }
    // ---- flush
    if (need_emb) {
        double* gC = bf.shared + SVGPFA_SHARED_HDR;
        double* gd = gC + (size_t)N * K;
        for (int idx = tid; idx < K * EM_TN; idx += EM_THREADS) {
            const int kk = idx / EM_TN, nn = idx - kk * EM_TN;
            if (n0 + nn < N) atomicAdd(gC + (size_t)(n0 + nn) * K + kk, dCs[kk * EM_TNS + nn]);
        }
        if (half == 0 && n < N) atomicAdd(gd + n, dd_acc);
    }
    const double tot = block_sum(t1, red);
    if (tid == 0) {
        const int slot = (blockIdx.y * gridDim.x + blockIdx.x) % SVGPFA_TERM1_SLOTS;
        atomicAdd(bf.term1_part + slot, tot);
    }
}

}  // namespace

extern "C" int svgpfa_quad_latent_fwd(const svgpfa_dims* dims, const svgpfa_buffers* buf, void* stream) {
    if (!dims || !buf || dims->Mmax > SVGPFA_MAX_M) return svgpfa_set_error(SVGPFA_E_ARG, "quad_latent_fwd", cudaSuccess);
    if (dims->R == 0 || dims->Q == 0) return SVGPFA_OK;
    const size_t smem = ql_smem_bytes(round_up(dims->Mmax, 4), false);
    cudaFuncSetAttribute(quad_latent_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    quad_latent_fwd_kernel<<<dim3(dims->R, dims->K), QL_TQ, smem, (cudaStream_t)stream>>>(*dims, *buf);
    SVGPFA_CHECK_LAUNCH("quad_latent_fwd");
    return SVGPFA_OK;
}

extern "C" int svgpfa_quad_latent_bwd(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, void* stream) {
    if (!dims || !buf || dims->Mmax > SVGPFA_MAX_M) return svgpfa_set_error(SVGPFA_E_ARG, "quad_latent_bwd", cudaSuccess);
    if (dims->R == 0) return SVGPFA_OK;
    const size_t smem = ql_smem_bytes(round_up(dims->Mmax, 4), true);
    if (smem > 227 * 1024) return svgpfa_set_error(SVGPFA_E_UNSUPPORTED, "quad_latent_bwd: shared memory", cudaSuccess);
    cudaFuncSetAttribute(quad_latent_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    quad_latent_bwd_kernel<<<dim3(dims->R, dims->K), QL_TQ, smem, (cudaStream_t)stream>>>(*dims, *buf, flags);
    SVGPFA_CHECK_LAUNCH("quad_latent_bwd");
    return SVGPFA_OK;
}

extern "C" int svgpfa_quad_embed_fwd_bwd(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, void* stream) {
    if (!dims || !buf) return svgpfa_set_error(SVGPFA_E_ARG, "quad_embed_fwd_bwd", cudaSuccess);
    if (dims->R == 0 || dims->Q == 0 || dims->N == 0) return SVGPFA_OK;
    const size_t smem = em_smem_bytes(dims->K);
    if (smem > 227 * 1024) return svgpfa_set_error(SVGPFA_E_UNSUPPORTED, "quad_embed: K too large for shared memory", cudaSuccess);
    cudaFuncSetAttribute(quad_embed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const int ntiles = (dims->N + EM_TN - 1) / EM_TN;
    const int qtiles = (dims->Q + EM_TQ - 1) / EM_TQ;
    const long nitems = (long)dims->R * qtiles;
    int dev = 0, nsm = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
    long workers = (long)nsm * 4 / ntiles;
    if (workers < 1) workers = 1;
    if (workers > nitems) workers = nitems;
    quad_embed_kernel<<<dim3(ntiles, (unsigned)workers), EM_THREADS, smem, (cudaStream_t)stream>>>(*dims, *buf, flags);
    SVGPFA_CHECK_LAUNCH("quad_embed_fwd_bwd");
    return SVGPFA_OK;
}

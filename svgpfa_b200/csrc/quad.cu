// Legendre-quadrature part of the expected log-likelihood (sm_100a, float64):
//   quad_latent_fwd_kernel   mu, var of every latent at every quadrature point        (R,Q,K)
//   quad_embed_kernel        h = C x + d, exp(mean + var/2), weighted integral, dC, dd, mubar, varbar
//   quad_latent_bwd_kernel   adjoints of the latent posterior: A_q, abar_q, dz_acc, dth_part
// Ktz (R,Q,M), Kzz^-1 Kzt (R,M,Q) and eLinkValues (R,Q,N) of the reference are never materialised
// (stats/kernelsMatricesStore.py:186-195, stats/svPosteriorOnLatents.py:185-216,
//  stats/svEmbedding.py:80-84, stats/expectedLogLikelihood.py:107-135,205-208).
#include "common.cuh"

bool svgpfa_try_quad_latent_big(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, bool bwd, cudaStream_t st);
bool svgpfa_try_quad_latent_mma(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, bool bwd,
                                cudaStream_t st);

bool svgpfa_try_quad_embed_mma(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, cudaStream_t st);

// The tensor-path kernels (quad_mma.cu) cover M <= 32 and K <= 39; the kernels in this file are the general path.
static bool use_mma_path() { return true; }

namespace {

// ======================================================================================
// latent posterior at quadrature points, one CTA (4 warps) per (trial, latent).
//   lane  <-> quadrature point of the current tile,
//   warp  <-> a PAIR of 4-row blocks (c, nb-1-c) of the triangular M x M factors, so that every warp does the
//             same number of multiply-adds in every triangular matrix-vector phase; when M <= 16 the spare
//             warps take further 32-point sub-tiles instead.
// Matrix entries are warp-uniform 16-byte shared-memory broadcasts, the per-point vectors (k, v, u, w) live in
// shared memory as [row][point] with an odd stride (conflict-free).  Nothing of size Q x M reaches HBM.
// ======================================================================================
constexpr int QL_THREADS = 128;
constexpr int QL_WARPS = QL_THREADS / 32;

struct QLGeom {
    int M, MP, nb, npair, npw, nqs, TQ, TQS;
};

__host__ __device__ inline QLGeom ql_geom(int M) {
    QLGeom g;
    g.M = M;
    g.MP = round_up(M, 4);
    g.nb = g.MP / 4;
    g.npair = (g.nb + 1) / 2;
    g.npw = g.npair < QL_WARPS ? g.npair : QL_WARPS;      // warps cooperating on one 32-point sub-tile
    g.nqs = QL_WARPS / g.npw;                             // 32-point sub-tiles per pass
    g.TQ = 32 * g.nqs;
    g.TQS = g.TQ + 1;
    return g;
}

__host__ __device__ inline size_t ql_smem_doubles(int Mmax, bool bwd) {
    // worst case over M <= Mmax of (mats + vectors); vectors shrink as MP grows only through TQ, so take both ends
    size_t best = 0;
    for (int M = 1; M <= Mmax; ++M) {
        const QLGeom g = ql_geom(M);
        const size_t n = (size_t)(bwd ? 4 : 2) * g.MP * g.MP + (size_t)(bwd ? 3 : 2) * g.MP * g.TQS + 2 * g.MP
                         + 4 * g.TQ + (size_t)QL_WARPS * g.TQ * 2;
        if (n > best) best = n;
    }
    return best;
}

struct QLSmem {
    double *LiT, *X, *Li, *XT;        // MP x MP (zero padded)
    double *ks, *vs, *us;             // MP x TQS   (ks doubles as w in the backward pass)
    double *al, *zs;                  // MP
    double *mb, *vb, *tt, *spare;     // TQ
    double *part;                     // QL_WARPS x TQ x 2 partial sums
};

__device__ __forceinline__ QLSmem ql_carve(double* sm, const QLGeom& g, bool bwd) {
    QLSmem s;
    double* p = sm;
    s.LiT = p; p += g.MP * g.MP;
    s.X = p; p += g.MP * g.MP;
    if (bwd) { s.Li = p; p += g.MP * g.MP; s.XT = p; p += g.MP * g.MP; } else { s.Li = s.XT = nullptr; }
    s.ks = p; p += g.MP * g.TQS;
    s.vs = p; p += g.MP * g.TQS;
    if (bwd) { s.us = p; p += g.MP * g.TQS; } else s.us = nullptr;
    s.al = p; p += g.MP;
    s.zs = p; p += g.MP;
    s.mb = p; p += g.TQ;
    s.vb = p; p += g.TQ;
    s.tt = p; p += g.TQ;
    s.spare = p; p += g.TQ;
    s.part = p;
    return s;
}

__device__ __forceinline__ void ql_load_mats(const QLSmem& s, const QLGeom& g, const svgpfa_dims& dm,
                                             const svgpfa_buffers& bf, const svgpfa_latent_desc& ds, int r, bool bwd) {
    const int M = g.M, MP = g.MP;
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

// One 4-row block of out = Mat * in for this lane's point.  Mt is Mat TRANSPOSED (Mt[j*MP + i] = Mat[i][j]);
// LOWER: Mat lower-triangular (j <= i), else upper-triangular (j >= i).
template <bool LOWER>
__device__ __forceinline__ void tri_block4(const double* __restrict__ Mt, const double* __restrict__ in, int MP, int M,
                                           int TQS, int col, int blk, double& a0, double& a1, double& a2, double& a3) {
    const int i0 = 4 * blk;
    a0 = a1 = a2 = a3 = 0.0;
    const int jb = LOWER ? 0 : i0;
    const int je = LOWER ? min(i0 + 4, M) : M;
    const double* mp = Mt + jb * MP + i0;
    const double* ip = in + jb * TQS + col;
#pragma unroll 4
    for (int j = jb; j < je; ++j) {
        const double x = *ip;
        const double2 m01 = *reinterpret_cast<const double2*>(mp);
        const double2 m23 = *reinterpret_cast<const double2*>(mp + 2);
        a0 = fma(m01.x, x, a0);
        a1 = fma(m01.y, x, a1);
        a2 = fma(m23.x, x, a2);
        a3 = fma(m23.y, x, a3);
        mp += MP;
        ip += TQS;
    }
}

// Runs f(blk) for every 4-row block owned by this warp: pairs (c, nb-1-c), c = pw, pw + npw, ...
template <class F>
__device__ __forceinline__ void for_my_blocks(const QLGeom& g, int pw, F f) {
    for (int c = pw; c < g.npair; c += g.npw) {
        f(c);
        const int c2 = g.nb - 1 - c;
        if (c2 != c) f(c2);
    }
}

__global__ void __launch_bounds__(QL_THREADS) quad_latent_fwd_kernel(svgpfa_dims dm, svgpfa_buffers bf) {
    extern __shared__ __align__(16) double sm[];
    __shared__ double etab[64];
    svgpfa_load_exp_tab64(etab);
    const int r = dm.r0 + blockIdx.x, k = blockIdx.y;
    const svgpfa_latent_desc ds = bf.desc[k];
    const QLGeom g = ql_geom(ds.M);
    const QLSmem s = ql_carve(sm, g, false);
    const KConst kc = make_kconst(ds, bf.theta, bf.kscale, k);
    ql_load_mats(s, g, dm, bf, ds, r, false);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int pw = warp % g.npw, qs = warp / g.npw;          // block-pair slot and 32-point sub-tile of this warp
    const bool warp_on = qs < g.nqs;
    const int col = qs * 32 + lane;
    const int M = g.M, MP = g.MP, TQS = g.TQS;
    __syncthreads();
    for (int q0 = 0; q0 < dm.Q; q0 += g.TQ) {
        // phase 0: kernel values, thread (pw, col) takes inducing points pw, pw + npw, ...
        double mu = 0.0;
        if (warp_on) {
            const int q = q0 + col;
            const double t = (q < dm.Q) ? bf.tq[(size_t)r * dm.Q + q] : 0.0;
            for (int j = pw; j < MP; j += g.npw) {
                const double kv = (q < dm.Q && j < M) ? kappa_val_t(kc, t - s.zs[j], etab) : 0.0;
                s.ks[j * TQS + col] = kv;
                mu = fma(kv, s.al[j], mu);
            }
        }
        __syncthreads();
        double vv = 0.0, uu = 0.0;
        if (warp_on) {
            for_my_blocks(g, pw, [&](int blk) {
                double a0, a1, a2, a3;
                tri_block4<true>(s.LiT, s.ks, MP, M, TQS, col, blk, a0, a1, a2, a3);       // v = Li k
                double* o = s.vs + 4 * blk * TQS + col;
                o[0] = a0; o[TQS] = a1; o[2 * TQS] = a2; o[3 * TQS] = a3;
                vv += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
            });
        }
        __syncthreads();
        if (warp_on) {
            for_my_blocks(g, pw, [&](int blk) {
                double a0, a1, a2, a3;
                tri_block4<false>(s.X, s.vs, MP, M, TQS, col, blk, a0, a1, a2, a3);        // u = X^T v
                uu += a0 * a0 + a1 * a1 + a2 * a2 + a3 * a3;
            });
            s.part[(pw * g.TQ + col) * 2 + 0] = mu;
            s.part[(pw * g.TQ + col) * 2 + 1] = uu - vv;
        }
        __syncthreads();
        if (warp_on && pw == 0) {
            const int q = q0 + col;
            if (q < dm.Q) {
                double m_ = 0.0, d_ = 0.0;
                for (int w = 0; w < g.npw; ++w) {
                    m_ += s.part[(w * g.TQ + col) * 2 + 0];
                    d_ += s.part[(w * g.TQ + col) * 2 + 1];
                }
                const size_t o = ((size_t)r * dm.Q + q) * dm.K + k;
                bf.mu_q[o] = m_;
                bf.var_q[o] = kc.s2 + d_;
            }
        }
        // the next pass's phase-0 writes (ks, part after 3 more barriers) cannot race with the reads above:
        // ks is only read in the v phase (two barriers back); part is rewritten after two further barriers.
    }
}

// BIG (M > 44): more 4x4 tiles of A than threads, every thread may own a second tile.
template <bool BIG>
__global__ void __launch_bounds__(QL_THREADS, BIG ? 1 : 4) quad_latent_bwd_kernel(svgpfa_dims dm, svgpfa_buffers bf, uint32_t flags) {
    extern __shared__ __align__(16) double sm[];
    __shared__ double red[32];
    __shared__ double etab[64];
    svgpfa_load_exp_tab64(etab);
    const int r = dm.r0 + blockIdx.x, k = blockIdx.y;
    const svgpfa_latent_desc ds = bf.desc[k];
    const QLGeom g = ql_geom(ds.M);
    const bool need_kz = flags & (SVGPFA_GRAD_KERNEL | SVGPFA_GRAD_INDLOCS);
    const QLSmem s = ql_carve(sm, g, true);
    const KConst kc = make_kconst(ds, bf.theta, bf.kscale, k);
    ql_load_mats(s, g, dm, bf, ds, r, true);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int pw = warp % g.npw, qs = warp / g.npw;
    const bool warp_on = qs < g.nqs;
    const int col = qs * 32 + lane;
    const int M = g.M, MP = g.MP, TQS = g.TQS;
    // ownership of the reduction A = sum_q varbar_q v v^T: 4x4 tiles of the lower triangle, G groups split q
    const int nt = g.nb, ntile = nt * (nt + 1) / 2;
    const int G = max(1, QL_THREADS / ntile);
    const int my_tile = tid % ntile, my_g = tid / ntile;
    const bool has_tile = tid < ntile * G;
    const int my_tile2 = tid + QL_THREADS;
    const bool has_tile2 = BIG && (G == 1) && (my_tile2 < ntile);
    auto decode = [](int t, int& ti, int& tj) {
        int row = 0;
        while (t >= row + 1) { t -= row + 1; ++row; }
        ti = row; tj = t;
    };
    int ti = 0, tj = 0, ti2 = 0, tj2 = 0;
    decode(my_tile, ti, tj);
    if (has_tile2) decode(my_tile2, ti2, tj2);
    double acc[16], acc2[BIG ? 16 : 1];
#pragma unroll
    for (int e = 0; e < 16; ++e) acc[e] = 0.0;
#pragma unroll
    for (int e = 0; e < (BIG ? 16 : 1); ++e) acc2[e] = 0.0;
    // per-lane running sums of abar_j / dz_j: lane e of warp (pw, qs) owns the e-th inducing point handled by that warp
    double ab_own = 0.0, dz_own = 0.0, th0 = 0.0, th1 = 0.0;
    const size_t part_stride = (size_t)dm.R * dm.Q * dm.K;
    __syncthreads();
    for (int q0 = 0; q0 < dm.Q; q0 += g.TQ) {
        // phase 0: adjoints of the statistics, kernel values, abar_j = sum_q mubar_q kappa(t_q, z_j)
        double t = 0.0, mbar = 0.0, vbar = 0.0;
        bool valid = false;
        if (warp_on) {
            const int q = q0 + col;
            valid = q < dm.Q;
            if (valid) {
                t = bf.tq[(size_t)r * dm.Q + q];
                const size_t o = ((size_t)r * dm.K + k) * dm.Q + q;          // [tile][r][k][q]: coalesced over q
                for (int p = 0; p < dm.n_ntiles; ++p) {
                    mbar += bf.mubar_part[p * part_stride + o];
                    vbar += bf.varbar_part[p * part_stride + o];
                }
            }
            if (pw == 0) { s.vb[col] = vbar; }
            int e = 0;
            for (int j = pw; j < MP; j += g.npw, ++e) {
                const double kv = (valid && j < M) ? kappa_val_t(kc, t - s.zs[j], etab) : 0.0;
                s.ks[j * TQS + col] = kv;
                const double sa = warp_sum(mbar * kv);
                if (lane == e) ab_own += sa;
            }
        }
        __syncthreads();
        if (warp_on) {
            for_my_blocks(g, pw, [&](int blk) {
                double a0, a1, a2, a3;
                tri_block4<true>(s.LiT, s.ks, MP, M, TQS, col, blk, a0, a1, a2, a3);       // v = Li k
                double* o = s.vs + 4 * blk * TQS + col;
                o[0] = a0; o[TQS] = a1; o[2 * TQS] = a2; o[3 * TQS] = a3;
            });
        }
        __syncthreads();
        if (need_kz) {
            if (warp_on) {
                for_my_blocks(g, pw, [&](int blk) {
                    double a0, a1, a2, a3;
                    tri_block4<false>(s.X, s.vs, MP, M, TQS, col, blk, a0, a1, a2, a3);    // u = X^T v
                    double* o = s.us + 4 * blk * TQS + col;
                    o[0] = a0; o[TQS] = a1; o[2 * TQS] = a2; o[3 * TQS] = a3;
                });
            }
            __syncthreads();
            if (warp_on) {
                for_my_blocks(g, pw, [&](int blk) {
                    double a0, a1, a2, a3;
                    tri_block4<true>(s.XT, s.us, MP, M, TQS, col, blk, a0, a1, a2, a3);    // w = X u - v  (into ks)
                    double* o = s.ks + 4 * blk * TQS + col;
                    const double* v = s.vs + 4 * blk * TQS + col;
                    o[0] = a0 - v[0]; o[TQS] = a1 - v[TQS]; o[2 * TQS] = a2 - v[2 * TQS]; o[3 * TQS] = a3 - v[3 * TQS];
                });
            }
            __syncthreads();
            if (warp_on) {
                int e = 0;
                for_my_blocks(g, pw, [&](int blk) {
                    double a[4];
                    tri_block4<false>(s.Li, s.ks, MP, M, TQS, col, blk, a[0], a[1], a[2], a[3]);   // Li^T w
#pragma unroll
                    for (int c = 0; c < 4; ++c, ++e) {
                        const int j = 4 * blk + c;
                        double gz = 0.0;
                        if (valid && j < M) {
                            const double kbar = 2.0 * vbar * a[c] + mbar * s.al[j];
                            double kv, dkd, d0, d1;
                            kappa_grad_t(kc, t - s.zs[j], etab, kv, dkd, d0, d1);
                            gz = -kbar * dkd;                        // d delta / d z = -1
                            th0 = fma(kbar, d0, th0);
                            th1 = fma(kbar, d1, th1);
                        }
                        const double sz = warp_sum(gz);
                        if (lane == e) dz_own += sz;
                    }
                });
            }
        }
        // phase 5: A += sum_q varbar_q v_q v_q^T over the points of this pass (vs is complete since the barrier
        // after the v phase; vb since the first barrier)
        if (has_tile) {
            for (int qq = my_g; qq < g.TQ; qq += G) {
                const double sv = s.vb[qq];
                double vi[4], vj[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    vi[e] = s.vs[(4 * ti + e) * TQS + qq] * sv;
                    vj[e] = s.vs[(4 * tj + e) * TQS + qq];
                }
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc[a * 4 + b] = fma(vi[a], vj[b], acc[a * 4 + b]);
            }
        }
        if (BIG && has_tile2) {
            for (int qq = 0; qq < g.TQ; ++qq) {
                const double sv = s.vb[qq];
                double vi[4], vj[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    vi[e] = s.vs[(4 * ti2 + e) * TQS + qq] * sv;
                    vj[e] = s.vs[(4 * tj2 + e) * TQS + qq];
                }
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) acc2[BIG ? a * 4 + b : 0] = fma(vi[a], vj[b], acc2[BIG ? a * 4 + b : 0]);
            }
        }
        __syncthreads();
    }
    // ---- write-out.  A: combine the G partial copies of every tile through shared memory (ks is free now)
    double* scratch = s.ks;
    const size_t mo = (size_t)r * dm.MM + ds.mmoff;
    for (int gi = 0; gi < G; ++gi) {
        if (has_tile && my_g == gi) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                if (gi == 0) scratch[my_tile * 16 + e] = acc[e];
                else scratch[my_tile * 16 + e] += acc[e];
            }
        }
        __syncthreads();
    }
    if (BIG && has_tile2) {
#pragma unroll
        for (int e = 0; e < 16; ++e) scratch[my_tile2 * 16 + e] = acc2[BIG ? e : 0];
    }
    __syncthreads();
    for (int idx = tid; idx < ntile * 16; idx += QL_THREADS) {
        const int tl = idx / 16, e = idx - tl * 16;
        int a_, b_;
        decode(tl, a_, b_);
        const int i = 4 * a_ + e / 4, j = 4 * b_ + (e & 3);
        if (i < M && j <= i) bf.A_q[mo + (size_t)i * M + j] = scratch[idx];
    }
    // abar / dz: the q sub-tiles (qs) hold partial sums for the same inducing points -> combine through smem
    double* comb = s.vs;                 // [2][nqs][MP]
    __syncthreads();
    if (warp_on) {
        // inducing points of this warp, in the order they were enumerated above
        int e = 0;
        for (int j = pw; j < MP; j += g.npw, ++e)
            if (lane == e) comb[(0 * g.nqs + qs) * MP + j] = ab_own;
        if (need_kz) {
            e = 0;
            for_my_blocks(g, pw, [&](int blk) {
                for (int c = 0; c < 4; ++c, ++e)
                    if (lane == e) comb[(1 * g.nqs + qs) * MP + 4 * blk + c] = dz_own;
            });
        }
    }
    __syncthreads();
    const size_t vo = (size_t)r * dm.KM + ds.moff;
    if (tid < M) {
        double sa = 0.0, sz = 0.0;
        for (int q = 0; q < g.nqs; ++q) {
            sa += comb[(0 * g.nqs + q) * MP + tid];
            if (need_kz) sz += comb[(1 * g.nqs + q) * MP + tid];
        }
        bf.abar_q[vo + tid] = sa;
        if (need_kz) atomicAdd(bf.dz_acc + vo + tid, sz);  // zeroed by the caller; the spike kernel adds concurrently
    }
    if (need_kz && (flags & SVGPFA_GRAD_KERNEL)) {
        const double s0 = block_sum(th0, red);
        const double s1 = block_sum(th1, red);
        if (tid == 0) {
            double* dth = bf.dth_part + (size_t)r * dm.TH + ds.thoff;
            atomicAdd(dth, s0);                           // zeroed by the caller; the spike kernel adds concurrently
            if (ds.nth > 1) atomicAdd(dth + 1, s1);
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

__global__ void __launch_bounds__(EM_THREADS, 3) quad_embed_kernel(svgpfa_dims dm, svgpfa_buffers bf, uint32_t flags) {
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
    const int nitems = (dm.rn ? dm.rn : dm.R) * qtiles;
    const size_t part_off = (size_t)tile * dm.R * Q * K;
    __syncthreads();
    for (int it = blockIdx.y; it < nitems; it += gridDim.y) {
        const int rl = it / qtiles, r = dm.r0 + rl, q0 = (it - rl * qtiles) * EM_TQ;
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
        //      (one output per thread-iteration; a 2-point register tile was measured slower: fewer active threads)
        if (need_lat) {
            for (int o = tid; o < EM_TQ * K; o += EM_THREADS) {
                const int kk = o / EM_TQ, qq = o - kk * EM_TQ;
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
                    const size_t oo = part_off + ((size_t)r * K + kk) * Q + q0 + qq;     // [tile][r][k][q]
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

// Post-fit read-out: embedding mean / variance and expected intensity at arbitrary times from the
// latent statistics, one thread per (trial, time, neuron); C is read through the read-only cache.
//   e_mean = mu C^T + d, e_var = var (C^T)^2 (svEmbedding.py:80-92), cif = exp(e_mean + e_var / 2)
//   (expectedLogLikelihood.py:62-73)
__global__ void __launch_bounds__(256) embed_predict_kernel(svgpfa_dims dm, svgpfa_buffers bf, double* __restrict__ e_mean,
                                                            double* __restrict__ e_var, double* __restrict__ cif) {
    const size_t total = (size_t)dm.R * dm.Q * dm.N;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t rq = idx / dm.N;
        const int n = (int)(idx - rq * dm.N);
        const double* mu = bf.mu_q + rq * dm.K;
        const double* var = bf.var_q + rq * dm.K;
        const double* c = bf.C + (size_t)n * dm.K;
        double m = bf.d[n], v = 0.0;
        for (int k = 0; k < dm.K; ++k) {
            const double ck = __ldg(c + k);
            m = fma(mu[k], ck, m);
            v = fma(var[k], ck * ck, v);
        }
        if (e_mean) e_mean[idx] = m;
        if (e_var) e_var[idx] = v;
        if (cif) cif[idx] = exp(fma(0.5, v, m));
    }
}

}  // namespace

extern "C" int svgpfa_embed_predict(const svgpfa_dims* dims, const svgpfa_buffers* buf, double* e_mean, double* e_var,
                                    double* cif, void* stream) {
    if (!dims || !buf || !buf->mu_q || !buf->var_q) return svgpfa_set_error(SVGPFA_E_ARG, "embed_predict", cudaSuccess);
    const size_t total = (size_t)dims->R * dims->Q * dims->N;
    if (total == 0) return SVGPFA_OK;
    size_t blocks = (total + 255) / 256;
    if (blocks > (size_t)svgpfa_sm_count() * 16) blocks = (size_t)svgpfa_sm_count() * 16;
    embed_predict_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(*dims, *buf, e_mean, e_var, cif);
    SVGPFA_CHECK_LAUNCH("embed_predict");
    return SVGPFA_OK;
}

extern "C" int svgpfa_quad_latent_fwd(const svgpfa_dims* dims, const svgpfa_buffers* buf, void* stream) {
    if (!dims || !buf || dims->Mmax > SVGPFA_MAX_M) return svgpfa_set_error(SVGPFA_E_ARG, "quad_latent_fwd", cudaSuccess);
    if (dims->R == 0 || dims->Q == 0) return SVGPFA_OK;
    if (use_mma_path() && (svgpfa_try_quad_latent_mma(dims, buf, 0, false, (cudaStream_t)stream) ||
                           svgpfa_try_quad_latent_big(dims, buf, 0, false, (cudaStream_t)stream))) {
        SVGPFA_CHECK_LAUNCH("quad_latent_fwd (mma)");
        return SVGPFA_OK;
    }
    const size_t smem = sizeof(double) * ql_smem_doubles(dims->Mmax, false);
    SVGPFA_ENSURE_SMEM(smem, quad_latent_fwd_kernel);
    quad_latent_fwd_kernel<<<dim3(svgpfa_ntrials(dims), dims->K), QL_THREADS, smem, (cudaStream_t)stream>>>(*dims, *buf);
    SVGPFA_CHECK_LAUNCH("quad_latent_fwd");
    return SVGPFA_OK;
}

// forward with V of every quadrature point still valid in buffers.v_q (same Z, theta as the call that wrote it)
extern "C" int svgpfa_quad_latent_fwd_cached(const svgpfa_dims* dims, const svgpfa_buffers* buf, void* stream) {
    if (!dims || !buf || dims->Mmax > SVGPFA_MAX_M) return svgpfa_set_error(SVGPFA_E_ARG, "quad_latent_fwd_cached", cudaSuccess);
    if (dims->R == 0 || dims->Q == 0) return SVGPFA_OK;
    if (buf->v_q && use_mma_path() && (svgpfa_try_quad_latent_mma(dims, buf, SVGPFA_REUSE_VQ, false, (cudaStream_t)stream) ||
                                       svgpfa_try_quad_latent_big(dims, buf, SVGPFA_REUSE_VQ, false, (cudaStream_t)stream))) {
        SVGPFA_CHECK_LAUNCH("quad_latent_fwd_cached (mma)");
        return SVGPFA_OK;
    }
    return svgpfa_quad_latent_fwd(dims, buf, stream);
}

extern "C" int svgpfa_quad_latent_bwd(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, void* stream) {
    if (!dims || !buf || dims->Mmax > SVGPFA_MAX_M) return svgpfa_set_error(SVGPFA_E_ARG, "quad_latent_bwd", cudaSuccess);
    if (dims->R == 0) return SVGPFA_OK;
    if (use_mma_path() && (svgpfa_try_quad_latent_mma(dims, buf, flags, true, (cudaStream_t)stream) ||
                           svgpfa_try_quad_latent_big(dims, buf, flags, true, (cudaStream_t)stream))) {
        SVGPFA_CHECK_LAUNCH("quad_latent_bwd (mma)");
        return SVGPFA_OK;
    }
    const size_t smem = sizeof(double) * ql_smem_doubles(dims->Mmax, true);
    if (smem > 227 * 1024) return svgpfa_set_error(SVGPFA_E_UNSUPPORTED, "quad_latent_bwd: shared memory", cudaSuccess);
    if (dims->Mmax > 44) {
        SVGPFA_ENSURE_SMEM(smem, quad_latent_bwd_kernel<true>);
        quad_latent_bwd_kernel<true><<<dim3(svgpfa_ntrials(dims), dims->K), QL_THREADS, smem, (cudaStream_t)stream>>>(*dims, *buf, flags);
    } else {
        SVGPFA_ENSURE_SMEM(smem, quad_latent_bwd_kernel<false>);
        quad_latent_bwd_kernel<false><<<dim3(svgpfa_ntrials(dims), dims->K), QL_THREADS, smem, (cudaStream_t)stream>>>(*dims, *buf, flags);
    }
    SVGPFA_CHECK_LAUNCH("quad_latent_bwd");
    return SVGPFA_OK;
}

extern "C" int svgpfa_quad_embed_fwd_bwd(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, void* stream) {
    if (!dims || !buf) return svgpfa_set_error(SVGPFA_E_ARG, "quad_embed_fwd_bwd", cudaSuccess);
    if (dims->R == 0 || dims->Q == 0 || dims->N == 0) return SVGPFA_OK;
    if (use_mma_path() && svgpfa_try_quad_embed_mma(dims, buf, flags, (cudaStream_t)stream)) {
        SVGPFA_CHECK_LAUNCH("quad_embed_fwd_bwd (mma)");
        return SVGPFA_OK;
    }
    const size_t smem = em_smem_bytes(dims->K);
    if (smem > 227 * 1024) return svgpfa_set_error(SVGPFA_E_UNSUPPORTED, "quad_embed: K too large for shared memory", cudaSuccess);
    SVGPFA_ENSURE_SMEM(smem, quad_embed_kernel);
    const int ntiles = (dims->N + EM_TN - 1) / EM_TN;
    const int qtiles = (dims->Q + EM_TQ - 1) / EM_TQ;
    const long nitems = (long)svgpfa_ntrials(dims) * qtiles;
    const int nsm = svgpfa_sm_count();
    int occ = 2;              // general path only (K > 39): the query runs per launch
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, quad_embed_kernel, EM_THREADS, smem);
    if (occ < 1) occ = 1;
    long workers = (long)nsm * occ / ntiles;           // one resident wave of persistent CTAs
    if (workers < 1) workers = 1;
    if (workers > nitems) workers = nitems;
    quad_embed_kernel<<<dim3(ntiles, (unsigned)workers), EM_THREADS, smem, (cudaStream_t)stream>>>(*dims, *buf, flags);
    SVGPFA_CHECK_LAUNCH("quad_embed_fwd_bwd");
    return SVGPFA_OK;
}

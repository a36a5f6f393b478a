// Latent posterior at the quadrature points and its adjoint for M <= 32, on the FP64 tensor path
// (mma.sync.m8n8k4.f64).  Same arithmetic as quad_latent_{fwd,bwd}_kernel in quad.cu (which stay as the
// M > 32 path); reference: stats/kernelsMatricesStore.py:186-195, stats/svPosteriorOnLatents.py:185-216.
//
// Why the tensor path for a 32 x 32 problem: on this part a DFMA holds the scheduler's issue port for two
// cycles, so the CUDA-core version of these triangular products is bound by instruction issue and by one
// shared-memory wavefront per FMA (profiles/README.md).  One m8n8k4 instruction carries 256 FMAs through the
// same FP64 units (measured: 37.1 TFLOP/s, identical to DFMA) for ONE issue slot and two operand registers.
//
// Mapping: CTA <-> (trial, latent); warp <-> a tile of 32 quadrature points, processed without block barriers.
//   V = Li K,  U = X^T V,  A += V diag(varbar) V^T,  W = X U - V,  Kv = Li^T W       (all 32 x 32 x 32 tiles)
// The kernel values are generated directly in B-fragment layout (no staging); V, U, W pass through a per-warp
// shared tile only to change from accumulator layout to B-fragment layout.  Leading dimensions (36) are chosen
// so that every fragment load/store is bank-conflict free.
#include <type_traits>

#include "common.cuh"

namespace {

constexpr int QM_MAX_WARPS = 4;               // measured on the 2000-trial shard: 7 warps (Q = 200: one 32-point tile per warp, one pass,
                                              // one CTA per SM in the adjoint) 4.76 ms against 3.87 ms for 4 warps (two CTAs per SM)
constexpr int QM_LDT = 36;                    // leading dimension of the per-warp 32-point tiles

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

__host__ __device__ inline size_t qm_smem_doubles(int MP, int nw, bool bwd) {
    const int ld = MP + 4;
    (void)bwd;                                    // one tile per warp in both kernels (see quad_latent_mma_kernel)
    return (size_t)2 * MP * ld + 2 * MP + (size_t)nw * (MP * QM_LDT + 3 * 32);
}

// Element-wise work (kernel evaluations, their derivatives, row sums) is written as ROLLED loops over a shared
// tile with lane <-> point or lane <-> inducing point, four independent evaluations per iteration (kappa_vals_n:
// the dependency chains are interleaved stage by stage); only the mma sequences are unrolled.  History
// (profiles/README.md): a first version kept the kernel values in fragment registers and unrolled everything
// (10 500 instructions, 23 % instruction-cache stalls); the round-1 forward kernel still generated K in B-fragment
// registers (32 inlined evaluations per pass: 16 % instruction-fetch stalls, 166 registers) and both kernels ran
// one evaluation chain at a time ("wait" was the top stall with two resident warps per scheduler).
// Work split over the quadrature points (round 2).  Round 1 gave every warp whole 32-point tiles: Q = 200 is 7 tiles
// (the last one 8 points wide) over 4 warps, i.e. two passes with the fourth warp idle in the second and 24 padded
// points -- 28 % of the tensor work wasted.  Now the points are cut into 8-point GROUPS (the n extent of one mma),
// the groups are dealt to the warps as evenly as possible (Q = 200: 25 groups -> 7, 6, 6, 6) and a warp covers its
// share in passes of 4 or 3 groups (7 -> 4 + 3, 6 -> 3 + 3): two instantiations of the pass body, no idle warp, at
// most two padded groups per warp.
// VC ("V cache"): V = Li K of every quadrature point goes to HBM in the forward kernel (buffers.v_q, R KM Q doubles) and
// is read back by the adjoint instead of being rebuilt there: the adjoint then needs neither the kernel values at the
// points (abar = sum_q mubar_q k_q = L sum_q mubar_q v_q, one M x M matrix-vector product per CTA at the end) nor the
// V product -- 20 of its 92 mma per 8 points and one of its two sets of kernel evaluations.
// VM = 0: no cache; 1: forward writes it, adjoint reads it; 2 (forward only): V is still valid for the current (Z, theta)
// -- every E-step closure after the first -- and the forward kernel reads it too: no kernel evaluation and no V product
// at all, mu_q = v_q . c (c = Li m), var_q = s2 - |v_q|^2 + |X^T v_q|^2.
template <int MT, bool BWD, int VM>
__global__ void __launch_bounds__(32 * QM_MAX_WARPS, (BWD || MT >= 4) ? 3 : 4)      // MT = 4: three CTAs fit an SM's shared memory anyway
quad_latent_mma_kernel(svgpfa_dims dm, svgpfa_buffers bf, uint32_t flags) {
    constexpr bool VC = VM >= 1, RV = !BWD && VM == 2, LOADV = VC && (BWD || RV);
    constexpr int MP = 8 * MT, KS = 2 * MT, LD = MP + 4, LDT = QM_LDT;
    extern __shared__ __align__(16) double sm[];
    __shared__ double red[32];
    __shared__ double etab[64];
    __shared__ double2 sctab[SVGPFA_SC_ENTRIES];
    svgpfa_load_exp_tab64(etab);
    const int r = dm.r0 + blockIdx.x, k = blockIdx.y;
    const svgpfa_latent_desc ds = bf.desc[k];
    const int M = ds.M;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    const int g = lane >> 2, tg = lane & 3;
    const bool need_kz = BWD && (flags & (SVGPFA_GRAD_KERNEL | SVGPFA_GRAD_INDLOCS));
    double* Lis = sm;                                   // [MP][LD]  Li
    double* Xs = Lis + MP * LD;                         // [MP][LD]  X; in the adjoint with need_kz: G = X X^T - I
    double* al = Xs + MP * LD;                          // [MP]
    double* zs = al + MP;                               // [MP]
    // ONE [MP][LDT] tile per warp holds, in turn, K, V, W = G V and Li^T W: every product reads the tile into fragment
    // registers completely before its result is written back (a __syncwarp on either side).  A second tile (round 1:
    // K / W beside V / Li^T W) cost 9 KB per warp and, with it, the third resident CTA per SM.
    constexpr int WSTRIDE = MP * LDT + 3 * 32;
    double* wbase = zs + MP + (size_t)warp * WSTRIDE;
    double* tileV = wbase;
    double* tileU = wbase;
    double* tileK = wbase;
    double* tt = wbase + MP * LDT;                      // [32] quadrature nodes of the pass
    double* mbs = tt + 32;                              // [32] mubar
    double* vbs = mbs + 32;                             // [32] varbar
    const KConst kc = make_kconst(ds, bf.theta, bf.kscale, k);
    if (kc.type == SVGPFA_KERNEL_PERIODIC) svgpfa_load_sincos_tab<1>(sctab);
    {
        const size_t mo = (size_t)r * dm.MM + ds.mmoff;
        // Li and X are first needed by the V product, after the kernel evaluations of the first pass: when the rows
        // are whole 16-byte chunks (M = MP, even offset) they are fetched with cp.async and awaited there, so the
        // 16 KB of loads run under ~1400 cycles of arithmetic (ncu: the staging was 12-15 % long-scoreboard stalls)
        // (the adjoint with the V cache and no kernel / inducing-point gradients -- an E-step closure -- reads neither)
        if (BWD && VC && !need_kz) {
        } else if (M == MP && (mo & 1) == 0) {
            const unsigned lis_s = (unsigned)__cvta_generic_to_shared(Lis), xs_s = (unsigned)__cvta_generic_to_shared(Xs);
            for (int c = tid; c < MP * (MP / 2); c += blockDim.x) {
                const int i = c / (MP / 2), jj = c - i * (MP / 2);
                const size_t go = mo + (size_t)i * M + 2 * jj;
                const unsigned so = (unsigned)(i * LD + 2 * jj) * 8u;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(lis_s + so), "l"(bf.Li + go) : "memory");
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(xs_s + so), "l"(bf.X + go) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        } else {
            for (int idx = tid; idx < MP * MP; idx += blockDim.x) {
                const int i = idx / MP, j = idx - i * MP;
                const bool in = (i < M) && (j < M);
                Lis[i * LD + j] = in ? bf.Li[mo + (size_t)i * M + j] : 0.0;
                Xs[i * LD + j] = in ? bf.X[mo + (size_t)i * M + j] : 0.0;
            }
        }
        const double* zg = bf.Z + (size_t)dm.R * ds.moff + (size_t)r * M;
        const size_t vo = (size_t)r * dm.KM + ds.moff;
        for (int i = tid; i < MP; i += blockDim.x) {
            zs[i] = (i < M) ? zg[i] : 0.0;
            al[i] = (i < M) ? (RV ? bf.c[vo + i] : bf.alpha[vo + i]) : 0.0;       // RV: mu = v . c
        }
    }
    __syncthreads();
    // persistent accumulators of the adjoint (BWD); lane j owns abar_j and the raw moments of dz_j, dtheta
    constexpr int NTA = MT * (MT + 1) / 2;
    double accA[BWD ? NTA : 1][2];
#pragma unroll
    for (int e = 0; e < (BWD ? NTA : 1); ++e) accA[e][0] = accA[e][1] = 0.0;
    double ab_own = 0.0, dz_raw = 0.0, th0_raw = 0.0, th1_raw = 0.0;
    const double zj_own = zs[lane < MP ? lane : 0], aj_own = al[lane < MP ? lane : 0];
    const size_t part_stride = (size_t)dm.R * dm.K * dm.Q;
    bool first_pass = true;
    if (LOADV && lane < M) {                                         // this warp's first V rows towards L2 (see prefetch_v)
        const int G0 = (dm.Q + 7) / 8, base0 = G0 / nw, rem0 = G0 - base0 * nw;
        const int q0 = 8 * (warp * base0 + (warp < rem0 ? warp : rem0));
        const double* vrow = bf.v_q + (((size_t)r * dm.KM + ds.moff + lane) * dm.Q + q0);
        asm volatile("prefetch.global.L2 [%0];" ::"l"(vrow));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(vrow + 16));
    }
    if (need_kz) {
        // G = X X^T - I in place of X (X itself is not needed by the adjoint): W = X X^T v - v = G v is ONE full product
        // (2 MT k-steps x NQT mma per row tile) instead of U = X^T V followed by W = X U - V (two triangular ones, a tile
        // round trip and a re-read of V between them).  Warp w computes the row tiles w, w + nw, ... into registers, all
        // warps meet, then the tiles are written over X.  (Done here, not under the first pass's kernel evaluations: the
        // block would be duplicated in both pass instantiations and its 2 MT^2 accumulators spill.)
        asm volatile("cp.async.wait_all;" ::: "memory");
        __syncthreads();
        first_pass = false;
        double gacc[MT][MT][2];                              // row tiles warp, warp + nw, ... (usually one)
#pragma unroll
        for (int o = 0; o < MT; ++o) {
            const int it = warp + o * nw;
            if (it < MT) {
#pragma unroll
                for (int jt = 0; jt < MT; ++jt) gacc[o][jt][0] = gacc[o][jt][1] = 0.0;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    const double a = Xs[(8 * it + g) * LD + 4 * ks + tg];
#pragma unroll
                    for (int jt = 0; jt < MT; ++jt)
                        dmma(gacc[o][jt][0], gacc[o][jt][1], a, Xs[(8 * jt + g) * LD + 4 * ks + tg]);
                }
            }
        }
        __syncthreads();
#pragma unroll
        for (int o = 0; o < MT; ++o) {
            const int it = warp + o * nw;
            if (it < MT) {
#pragma unroll
                for (int jt = 0; jt < MT; ++jt) {
                    const int i = 8 * it + g, j = 8 * jt + 2 * tg;
                    *reinterpret_cast<double2*>(Xs + i * LD + j) =
                        make_double2(gacc[o][jt][0] - (i == j ? 1.0 : 0.0), gacc[o][jt][1] - (i == j + 1 ? 1.0 : 0.0));
                }
            }
        }
        __syncthreads();
    }

    // ---- one pass over NQT groups (8 NQT <= 32 points) starting at point qbase
    // L2 prefetch of the V rows of a pass starting at point qb (V cache, adjoint): lane <-> row, two 128-byte lines each
    auto prefetch_v = [&](int qb) {
        if (LOADV && qb >= 0 && lane < M) {
            const double* vrow = bf.v_q + (((size_t)r * dm.KM + ds.moff + lane) * dm.Q + qb);
            asm volatile("prefetch.global.L2 [%0];" ::"l"(vrow));
            asm volatile("prefetch.global.L2 [%0];" ::"l"(vrow + 16));
        }
    };
    // t_pref: this lane's quadrature node of the pass, fetched one pass ahead (ncu: the first use of a node loaded at the
    // top of its own pass was the kernel's largest long-scoreboard stall)
    auto run_pass = [&](auto nqt_tag, const int qbase, const int qend, const int next_qbase, double& t_pref) {
        constexpr int NQT = decltype(nqt_tag)::value;
        constexpr int NPT = 8 * NQT;
        const int q_lane = qbase + lane;
        if (LOADV) {
            // the tile receives V straight from HBM: 16-byte asynchronous copies, two rows per instruction, zero fill for
            // rows >= M and points past the pass; all of them in flight at once, awaited below together with the
            // partial sums -- and the next pass's rows are pulled into L2 meanwhile
            const double* vg = bf.v_q + ((size_t)r * dm.KM + ds.moff) * dm.Q;
            const unsigned tile_s = (unsigned)__cvta_generic_to_shared(tileV);
            const int half = lane >> 4, cl = lane & 15, col = qbase + 2 * cl;
            const bool colok = 2 * cl < NPT && col < qend;
#pragma unroll 4
            for (int j0 = 0; j0 < MP; j0 += 2) {
                const int j = j0 + half;
                const bool ok = colok && j < M;
                const double* src = vg + (ok ? (size_t)j * dm.Q + col : 0);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tile_s + (unsigned)(j * LDT + 2 * cl) * 8u), "l"(src),
                             "r"(ok ? 16 : 0) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            prefetch_v(next_qbase);
        }
        const bool valid = lane < NPT && q_lane < qend;       // qend: end of this warp's share (a padded group belongs
                                                              // to the next warp) and of the trial's points
        const double t_lane = valid ? t_pref : 0.0;
        if (next_qbase >= 0 && next_qbase + lane < qend) t_pref = bf.tq[(size_t)r * dm.Q + next_qbase + lane];
        tt[lane] = t_lane;
        if (BWD && valid) {                          // the partial sums are read after the kernel evaluations (below):
            const size_t o = ((size_t)r * dm.K + k) * dm.Q + q_lane;      // pull them into L1 meanwhile
            for (int p = 0; p < dm.n_ntiles; ++p) {
                asm volatile("prefetch.global.L1 [%0];" ::"l"(bf.mubar_part + p * part_stride + o));
                asm volatile("prefetch.global.L1 [%0];" ::"l"(bf.varbar_part + p * part_stride + o));
            }
        }
        // ---- kernel values K[j][q], lane <-> point, four inducing points per iteration -> tileK.
        //      FWD: mu_q = k_q . alpha falls out of the same loop without any cross-lane reduction.
        //      BWD with the V cache: the tile receives V straight from HBM (rows >= M and points past the pass are zero).
        double mu_lane = 0.0, vv_lane = 0.0;
        if (LOADV) {
            // (V is on its way into the tile, see the top of the pass)
        } else {
#pragma unroll 1
            for (int j0 = 0; j0 < MP; j0 += 4) {
                double dl[4], kv[4], qq[4], s2x[4];
#pragma unroll
                for (int e = 0; e < 4; ++e) dl[e] = t_lane - zs[j0 + e];
                kappa_vals_n<4>(kc, dl, etab, sctab, kv, qq, s2x);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const double v = (valid && j0 + e < M) ? kv[e] : 0.0;
                    tileK[(j0 + e) * LDT + lane] = v;
                    if (!BWD) mu_lane = fma(v, al[j0 + e], mu_lane);
                }
            }
        }
        if (BWD) {
            double mbar = 0.0, vbar = 0.0;
            if (valid) {
                const size_t o = ((size_t)r * dm.K + k) * dm.Q + q_lane;
                for (int p = 0; p < dm.n_ntiles; ++p) {
                    mbar += bf.mubar_part[p * part_stride + o];
                    vbar += bf.varbar_part[p * part_stride + o];
                }
            }
            mbs[lane] = mbar;
            vbs[lane] = vbar;
            if (VC) asm volatile("cp.async.wait_all;" ::: "memory");
        } else {
            if (RV) {                                 // mu_q = v_q . c and |v_q|^2, lane <-> point
                asm volatile("cp.async.wait_all;" ::: "memory");
                __syncwarp();
                double m0 = 0.0, m1 = 0.0, w0 = 0.0, w1 = 0.0;
#pragma unroll 4
                for (int j = 0; j < MP; j += 2) {
                    const double x0 = tileV[j * LDT + lane], x1 = tileV[(j + 1) * LDT + lane];
                    m0 = fma(x0, al[j], m0);
                    m1 = fma(x1, al[j + 1], m1);
                    w0 = fma(x0, x0, w0);
                    w1 = fma(x1, x1, w1);
                }
                mu_lane = m0 + m1;
                vv_lane = w0 + w1;
            }
            if (valid) bf.mu_q[((size_t)r * dm.Q + q_lane) * dm.K + k] = mu_lane;
        }
        __syncwarp();
        if (BWD && lane < MP) {
            // abar_j += sum_q mubar_q K[j][q], lane <-> inducing point, skewed column order (conflict free); the
            // columns of a 3-group pass beyond 24 hold zeros
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll 2
            for (int c = 0; c < 32; c += 4) {
                const int q0 = (c + lane) & 31, q1 = (c + 1 + lane) & 31, q2 = (c + 2 + lane) & 31, q3 = (c + 3 + lane) & 31;
                s0 = fma(mbs[q0], tileU[lane * LDT + q0], s0);
                s1 = fma(mbs[q1], tileU[lane * LDT + q1], s1);
                s2 = fma(mbs[q2], tileU[lane * LDT + q2], s2);
                s3 = fma(mbs[q3], tileU[lane * LDT + q3], s3);
            }
            ab_own += (s0 + s1) + (s2 + s3);
        }
        if (first_pass) {                         // Li, X of the cp.async staging become visible to the whole CTA
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncthreads();
            first_pass = false;
        }
        // ---- V = Li K      v[rt][qt] = V[8 rt + g][8 qt + 2 tg + {0,1}]     (BWD with the V cache: already in the tile)
        double v[MT][NQT][2];
        if (!LOADV) {
#pragma unroll
            for (int rt = 0; rt < MT; ++rt)
#pragma unroll
                for (int qt = 0; qt < NQT; ++qt) v[rt][qt][0] = v[rt][qt][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                double b[NQT];
#pragma unroll
                for (int qt = 0; qt < NQT; ++qt) b[qt] = tileK[(4 * ks + tg) * LDT + 8 * qt + g];
#pragma unroll
                for (int rt = ks / 2; rt < MT; ++rt) {               // Li lower-triangular: k-step ks feeds row tiles >= ks/2
                    const double a = Lis[(8 * rt + g) * LD + 4 * ks + tg];
#pragma unroll
                    for (int qt = 0; qt < NQT; ++qt) dmma(v[rt][qt][0], v[rt][qt][1], a, b[qt]);
                }
            }
        }
        if (!BWD && VM == 1) {                                       // the forward kernel feeds the cache from the fragments
            double* vg = bf.v_q + ((size_t)r * dm.KM + ds.moff) * dm.Q;
#pragma unroll
            for (int rt = 0; rt < MT; ++rt)
#pragma unroll
                for (int qt = 0; qt < NQT; ++qt) {
                    const int row = 8 * rt + g, col = qbase + 8 * qt + 2 * tg;       // Q is even: pairs never straddle qend
                    if (row < M && col < qend)
                        __stcs(reinterpret_cast<double2*>(vg + (size_t)row * dm.Q + col), make_double2(v[rt][qt][0], v[rt][qt][1]));
                }
        }
        if (!BWD) {
            double vv[NQT][2];
            if (!RV) {
#pragma unroll
                for (int qt = 0; qt < NQT; ++qt) {
                    vv[qt][0] = vv[qt][1] = 0.0;
#pragma unroll
                    for (int rt = 0; rt < MT; ++rt) {
                        vv[qt][0] = fma(v[rt][qt][0], v[rt][qt][0], vv[qt][0]);
                        vv[qt][1] = fma(v[rt][qt][1], v[rt][qt][1], vv[qt][1]);
                    }
                }
                __syncwarp();                                        // every lane is done reading K (same tile)
#pragma unroll
                for (int rt = 0; rt < MT; ++rt)
#pragma unroll
                    for (int qt = 0; qt < NQT; ++qt)
                        *reinterpret_cast<double2*>(tileV + (8 * rt + g) * LDT + 8 * qt + 2 * tg) = make_double2(v[rt][qt][0], v[rt][qt][1]);
                __syncwarp();
            }
            // ---- U = X^T V     u[jt][qt] = U[8 jt + g][8 qt + 2 tg + {0,1}]
            double u[MT][NQT][2];
#pragma unroll
            for (int jt = 0; jt < MT; ++jt)
#pragma unroll
                for (int qt = 0; qt < NQT; ++qt) u[jt][qt][0] = u[jt][qt][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                double b[NQT];
#pragma unroll
                for (int qt = 0; qt < NQT; ++qt) b[qt] = tileV[(4 * ks + tg) * LDT + 8 * qt + g];
#pragma unroll
                for (int jt = 0; jt <= ks / 2 && jt < MT; ++jt) {    // X^T upper-triangular: k-step ks feeds row tiles <= ks/2
                    const double a = Xs[(4 * ks + tg) * LD + 8 * jt + g];
#pragma unroll
                    for (int qt = 0; qt < NQT; ++qt) dmma(u[jt][qt][0], u[jt][qt][1], a, b[qt]);
                }
            }
            // var = s2 - ||v||^2 + ||u||^2 : column sums over the 8 row groups
#pragma unroll
            for (int qt = 0; qt < NQT; ++qt) {
                double d0, d1;
                if (RV) {                             // |v|^2 of columns 8 qt + 2 tg + {0, 1}: the full sums, subtracted once
                    d0 = -__shfl_sync(0xffffffffu, vv_lane, 8 * qt + 2 * tg);
                    d1 = -__shfl_sync(0xffffffffu, vv_lane, 8 * qt + 2 * tg + 1);
                    if (g != 0) d0 = d1 = 0.0;
                } else {
                    d0 = -vv[qt][0];
                    d1 = -vv[qt][1];
                }
#pragma unroll
                for (int jt = 0; jt < MT; ++jt) {
                    d0 = fma(u[jt][qt][0], u[jt][qt][0], d0);
                    d1 = fma(u[jt][qt][1], u[jt][qt][1], d1);
                }
#pragma unroll
                for (int o = 4; o < 32; o <<= 1) {
                    d0 += __shfl_xor_sync(0xffffffffu, d0, o);
                    d1 += __shfl_xor_sync(0xffffffffu, d1, o);
                }
                if (g == 0) {                                      // lanes 0..3 hold columns 8 qt + 2 tg + {0,1}
                    const int q = qbase + 8 * qt + 2 * tg;
                    if (q < qend) bf.var_q[((size_t)r * dm.Q + q) * dm.K + k] = kc.s2 + d0;
                    if (q + 1 < qend) bf.var_q[((size_t)r * dm.Q + q + 1) * dm.K + k] = kc.s2 + d1;
                }
            }
        } else {
            if (!VC) {
                __syncwarp();                                        // every lane is done reading K (abar loop, V product)
#pragma unroll
                for (int rt = 0; rt < MT; ++rt)
#pragma unroll
                    for (int qt = 0; qt < NQT; ++qt)
                        *reinterpret_cast<double2*>(tileV + (8 * rt + g) * LDT + 8 * qt + 2 * tg) = make_double2(v[rt][qt][0], v[rt][qt][1]);
                __syncwarp();
            }
            // ---- A += V diag(varbar) V^T over the points of the pass: k-step = 4 points, A/B fragments from the V tile
#pragma unroll
            for (int ks = 0; ks < 2 * NQT; ++ks) {
                const double sv = vbs[4 * ks + tg];
                double av[MT];
#pragma unroll
                for (int it = 0; it < MT; ++it) av[it] = tileV[(8 * it + g) * LDT + 4 * ks + tg];
#pragma unroll
                for (int it = 0; it < MT; ++it)
#pragma unroll
                    for (int jt = 0; jt <= it; ++jt)
                        dmma(accA[it * (it + 1) / 2 + jt][0], accA[it * (it + 1) / 2 + jt][1], av[it] * sv, av[jt]);
            }
            if (need_kz) {
                // ---- W = G V   (G = X X^T - I, symmetric, full)
                double w[MT][NQT][2];
#pragma unroll
                for (int it = 0; it < MT; ++it)
#pragma unroll
                    for (int qt = 0; qt < NQT; ++qt) w[it][qt][0] = w[it][qt][1] = 0.0;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    double b[NQT];
#pragma unroll
                    for (int qt = 0; qt < NQT; ++qt) b[qt] = tileV[(4 * ks + tg) * LDT + 8 * qt + g];
#pragma unroll
                    for (int it = 0; it < MT; ++it) {
                        const double a = Xs[(8 * it + g) * LD + 4 * ks + tg];
#pragma unroll
                        for (int qt = 0; qt < NQT; ++qt) dmma(w[it][qt][0], w[it][qt][1], a, b[qt]);
                    }
                }
                __syncwarp();                                        // V fully consumed by every lane (SYRK, W product)
#pragma unroll
                for (int it = 0; it < MT; ++it)
#pragma unroll
                    for (int qt = 0; qt < NQT; ++qt)
                        *reinterpret_cast<double2*>(tileU + (8 * it + g) * LDT + 8 * qt + 2 * tg) = make_double2(w[it][qt][0], w[it][qt][1]);
                __syncwarp();
                // ---- Kv = Li^T W -> tileV (V is no longer needed once every lane is past the W product)
#pragma unroll
                for (int jt = 0; jt < MT; ++jt)
#pragma unroll
                    for (int qt = 0; qt < NQT; ++qt) w[jt][qt][0] = w[jt][qt][1] = 0.0;
#pragma unroll
                for (int ks = 0; ks < KS; ++ks) {
                    double b[NQT];
#pragma unroll
                    for (int qt = 0; qt < NQT; ++qt) b[qt] = tileU[(4 * ks + tg) * LDT + 8 * qt + g];
#pragma unroll
                    for (int jt = 0; jt <= ks / 2 && jt < MT; ++jt) {    // Li^T upper-triangular
                        const double a = Lis[(4 * ks + tg) * LD + 8 * jt + g];
#pragma unroll
                        for (int qt = 0; qt < NQT; ++qt) dmma(w[jt][qt][0], w[jt][qt][1], a, b[qt]);
                    }
                }
                __syncwarp();                                        // W fully consumed by every lane
#pragma unroll
                for (int jt = 0; jt < MT; ++jt)
#pragma unroll
                    for (int qt = 0; qt < NQT; ++qt)
                        *reinterpret_cast<double2*>(tileV + (8 * jt + g) * LDT + 8 * qt + 2 * tg) = make_double2(w[jt][qt][0], w[jt][qt][1]);
                __syncwarp();
                // ---- kbar = 2 varbar Kv + mubar alpha and its products with dkappa; lane <-> inducing point, four
                //      points per iteration, the NPT points of the pass in skewed order.  Raw moments only:
                //      h = kbar kappa, dz += h (delta | sin 2x), th0 += h (delta^2 | sin^2), th1 += h sin 2x delta;
                //      the constants dd, dl, dp are applied once at the end.
                if (lane < M) {
#pragma unroll 1
                    for (int c = 0; c < NPT; c += 4) {
                        double dl[4], kv[4], qq[4], s2x[4];
                        int qi[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            int q = c + e + lane;
                            if (q >= NPT) q -= NPT;
                            if (q >= NPT) q -= NPT;
                            qi[e] = q;
                            dl[e] = tt[q] - zj_own;
                        }
                        kappa_vals_n<4>(kc, dl, etab, sctab, kv, qq, s2x);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const double kbar = fma(2.0 * vbs[qi[e]], tileV[lane * LDT + qi[e]], mbs[qi[e]] * aj_own);
                            const double h = kbar * kv[e];
                            if (kc.type == SVGPFA_KERNEL_EXPQUAD) {
                                dz_raw = fma(h, dl[e], dz_raw);
                                th0_raw = fma(h, qq[e], th0_raw);
                            } else {
                                const double hs = h * s2x[e];
                                dz_raw += hs;
                                th0_raw = fma(h, qq[e], th0_raw);
                                th1_raw = fma(hs, dl[e], th1_raw);
                            }
                        }
                    }
                }
            }
        }
        __syncwarp();                                                // tiles are reused by the next pass
    };

    // ---- this warp's share of the 8-point groups and its passes (see above); every warp gets at least one group
    {
        const int G = (dm.Q + 7) / 8;
        const int base = G / nw, rem = G - base * nw;
        const int share = base + (warp < rem ? 1 : 0);
        int gstart = warp * base + (warp < rem ? warp : rem);
        const int qend = min(dm.Q, 8 * (gstart + share));
        const int np = (share + 3) / 4;
        int n4 = share - 3 * np;                                     // passes of four groups; the rest take three
        if (n4 < 0) n4 = 0;
        double t_pref = 8 * gstart + lane < qend ? bf.tq[(size_t)r * dm.Q + 8 * gstart + lane] : 0.0;
        for (int p = 0; p < np; ++p) {
            const int step = p < n4 ? 4 : 3, nxt = p + 1 < np ? 8 * (gstart + step) : -1;
            if (p < n4) run_pass(std::integral_constant<int, 4>{}, 8 * gstart, qend, nxt, t_pref);
            else run_pass(std::integral_constant<int, 3>{}, 8 * gstart, qend, nxt, t_pref);
            gstart += step;
        }
    }
    if (!BWD) return;
    // d delta / d z = -1;  dkappa/ddelta = kappa (delta | sin 2x) dd,  dkappa/dtheta0 = kappa (delta^2 | sin^2) dl,
    // dkappa/dtheta1 = kappa sin 2x delta dp  (kappa_grad in common.cuh); kappa's scale^2 is inside kv
    const double dz_own = -kc.dd * dz_raw, th0 = kc.dl * th0_raw, th1 = kc.dp * th1_raw;
    // ---- combine the warps through their (now idle) V tiles
    __syncthreads();
#pragma unroll
    for (int it = 0; it < MT; ++it)
#pragma unroll
        for (int jt = 0; jt <= it; ++jt) {
            const int tl = it * (it + 1) / 2 + jt;
            tileV[tl * 64 + g * 8 + 2 * tg] = accA[tl][0];
            tileV[tl * 64 + g * 8 + 2 * tg + 1] = accA[tl][1];
        }
    if (lane < MP) {
        tileV[NTA * 64 + lane] = ab_own;
        tileV[NTA * 64 + MP + lane] = dz_own;
    }
    __syncthreads();
    const double* t0p = zs + MP;                                     // warp 0's region
    const size_t mo = (size_t)r * dm.MM + ds.mmoff;
    for (int idx = tid; idx < NTA * 64; idx += blockDim.x) {
        double s = 0.0;
        for (int w = 0; w < nw; ++w) s += t0p[(size_t)w * WSTRIDE + idx];
        const int tl = idx / 64, e = idx - tl * 64;
        int it = 0, rem = tl;
        while (rem >= it + 1) { rem -= it + 1; ++it; }
        const int i = 8 * it + e / 8, j = 8 * rem + (e & 7);
        if (i < M && j <= i) bf.A_q[mo + (size_t)i * M + j] = s;
    }
    const size_t vo = (size_t)r * dm.KM + ds.moff;
    if (VC) {                                                        // L over Li (no longer needed) for abar = L vm below
        for (int idx = tid; idx < MP * MP; idx += blockDim.x) {
            const int i = idx / MP, j = idx - i * MP;
            Lis[i * LD + j] = (i < M && j <= i) ? bf.L[mo + (size_t)i * M + j] : 0.0;
        }
    }
    if (tid < MP) {
        double sa = 0.0, sz = 0.0;
        for (int w = 0; w < nw; ++w) {
            sa += t0p[(size_t)w * WSTRIDE + NTA * 64 + tid];
            sz += t0p[(size_t)w * WSTRIDE + NTA * 64 + MP + tid];
        }
        if (VC) al[tid] = sa;                                        // vm = sum_q mubar_q v_q
        else if (tid < M) bf.abar_q[vo + tid] = sa;
        if (need_kz && tid < M) atomicAdd(bf.dz_acc + vo + tid, sz);  // zeroed by the caller; the spike kernel adds too
    }
    if (VC) {
        __syncthreads();
        if (tid < M) {
            double s0 = 0.0, s1 = 0.0;
#pragma unroll 4
            for (int p = 0; p < MP; p += 2) {
                s0 = fma(Lis[tid * LD + p], al[p], s0);
                s1 = fma(Lis[tid * LD + p + 1], al[p + 1], s1);
            }
            bf.abar_q[vo + tid] = s0 + s1;
        }
    }
    if (need_kz && (flags & SVGPFA_GRAD_KERNEL)) {
        const double s0 = block_sum(th0, red);
        const double s1 = block_sum(th1, red);
        if (tid == 0) {
            double* dth = bf.dth_part + (size_t)r * dm.TH + ds.thoff;
            atomicAdd(dth, s0);                                      // zeroed by the caller; the spike kernel adds too
            if (ds.nth > 1) atomicAdd(dth + 1, s1);
        }
    }
}

template <int MT, bool BWD, int VC>
void launch_qm_vc(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, cudaStream_t st) {
    const int G = (dims->Q + 7) / 8;                    // 8-point groups
    int nw = (G + 3) / 4;                               // a warp takes up to four groups before another one is added
    int cap = dims->quad_warps > 0 ? dims->quad_warps : QM_MAX_WARPS;
    if (cap > QM_MAX_WARPS) cap = QM_MAX_WARPS;
    if (nw > cap) nw = cap;
    if (nw < 1) nw = 1;
    const size_t smem = sizeof(double) * qm_smem_doubles(8 * MT, nw, BWD);
    SVGPFA_ENSURE_SMEM(sizeof(double) * qm_smem_doubles(8 * MT, QM_MAX_WARPS, BWD), quad_latent_mma_kernel<MT, BWD, VC>);
    quad_latent_mma_kernel<MT, BWD, VC><<<dim3(svgpfa_ntrials(dims), dims->K), 32 * nw, smem, st>>>(*dims, *buf, flags);
}

template <int MT, bool BWD>
void launch_qm(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, cudaStream_t st) {
    // the V cache: buffers.v_q given, pairs of points 16-byte aligned (Q even)
    if (buf->v_q && (dims->Q & 1) == 0) {
        if (!BWD && (flags & SVGPFA_REUSE_VQ)) launch_qm_vc<MT, false, 2>(dims, buf, flags, st);
        else launch_qm_vc<MT, BWD, 1>(dims, buf, flags, st);
    } else launch_qm_vc<MT, BWD, 0>(dims, buf, flags, st);
}

// ======================================================================================
// embedding + exp link + integral on the tensor path.  Persistent CTA (8 warps) per (128-neuron tile, worker);
// items = (trial, 16 quadrature points).  Three small GEMMs per item share a shared-memory tile of G = -w exp(.):
//   A:  H = Mu C^T + d,  Sg = Var (C^2)^T           (16 x 128 x K)    -> G
//   B:  mubar = G C,  varbar = 1/2 G C^2             (16 x K x 128)    -> per-tile partials in HBM
//   C:  dC += G^T Mu + C o (G^T Var),  dd += G^T 1    (128 x K x 16)    -> accumulators in registers across items
// The ones-column trick: column K of the (zero-padded) Mu operand is 1, so dd falls out of GEMM C for free, while
// row K of C^T is 0 so GEMM A is unaffected.  Same arithmetic as quad_embed_kernel in quad.cu (reference:
// stats/svEmbedding.py:80-84, stats/expectedLogLikelihood.py:107-135,205-208).
// ======================================================================================
constexpr int EMM_TN = SVGPFA_EMBED_TN;       // 128
constexpr int EMM_TNS = EMM_TN + 4;           // 132 = 4 mod 16: conflict-free fragment loads
constexpr int EMM_TQ = 16;
constexpr int EMM_LDQ = 20;                   // leading dimension of the [k][q] statistics tiles
constexpr int EMM_THREADS = 256;

__host__ __device__ inline int emm_kp(int K) { return (K + 1 + 7) / 8 * 8; }       // K + ones column, padded to 8
__host__ __device__ inline size_t emm_smem_doubles(int K) {
    const int KP = emm_kp(K);
    return (size_t)2 * KP * EMM_TNS + (size_t)EMM_TQ * EMM_TNS + (size_t)2 * KP * EMM_LDQ + EMM_TQ + EMM_TN;
}

template <int KT>      // KT = KP / 8 k-tiles
__global__ void __launch_bounds__(EMM_THREADS) quad_embed_mma_kernel(svgpfa_dims dm, svgpfa_buffers bf, uint32_t flags) {
    constexpr int KP = 8 * KT, KS4 = KP / 4, TNS = EMM_TNS, LDQ = EMM_LDQ;
    extern __shared__ __align__(16) double sm[];
    __shared__ double red[32];
    const int K = dm.K, N = dm.N, Q = dm.Q;
    double* CT = sm;                               // [KP][TNS]   C^T, rows >= K are zero
    double* CT2 = CT + (size_t)KP * TNS;           // [KP][TNS]   (C^T)^2 (keeps a DMUL and two selects out of GEMM B's loop)
    double* Gs = CT2 + (size_t)KP * TNS;           // [TQ][TNS]
    double* muT = Gs + (size_t)EMM_TQ * TNS;       // [KP][LDQ]   row K = ones
    double* varT = muT + (size_t)KP * LDQ;         // [KP][LDQ]
    double* ws = varT + (size_t)KP * LDQ;          // [TQ]
    double* dvec = ws + EMM_TQ;                    // [TN]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tg = lane & 3;
    const int tile = blockIdx.x, n0 = tile * EMM_TN;
    const bool need_emb = flags & SVGPFA_GRAD_EMBEDDING;
    const bool need_lat = flags & (SVGPFA_GRAD_POSTERIOR | SVGPFA_GRAD_KERNEL | SVGPFA_GRAD_INDLOCS);
    for (int idx = tid; idx < KP * EMM_TN; idx += EMM_THREADS) {
        const int kk = idx / EMM_TN, nn = idx - kk * EMM_TN;
        const double cval = (kk < K && n0 + nn < N) ? bf.C[(size_t)(n0 + nn) * K + kk] : 0.0;
        CT[kk * TNS + nn] = cval;
        CT2[kk * TNS + nn] = cval * cval;
    }
    if (tid < EMM_TN) dvec[tid] = (n0 + tid < N) ? bf.d[n0 + tid] : 0.0;
    // GEMM C accumulators: warp owns neuron tiles {2 warp, 2 warp + 1} x all k-tiles, for Mu and for Var
    double cm[2][KT][2], cv[2][KT][2];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < KT; ++b) cm[a][b][0] = cm[a][b][1] = cv[a][b][0] = cv[a][b][1] = 0.0;
    double t1 = 0.0;
    const int qtiles = (Q + EMM_TQ - 1) / EMM_TQ;
    const int nitems = (dm.rn ? dm.rn : dm.R) * qtiles;
    const size_t part_off = (size_t)tile * dm.R * K * Q;
    __syncthreads();
    // The statistics of the NEXT item are fetched into registers while the three products of the current one run
    // (ncu: the in-loop global loads were 9 % of the stall samples, all threads waiting at the barrier behind them).
    constexpr int NPF = (EMM_TQ * KP + EMM_THREADS - 1) / EMM_THREADS;
    double pf_mu[NPF], pf_var[NPF], pf_w = 0.0;
    auto fetch = [&](int it) {
        const int rl = it / qtiles, r = dm.r0 + rl, q0 = (it - rl * qtiles) * EMM_TQ;
#pragma unroll
        for (int e = 0; e < NPF; ++e) {
            const int idx = tid + e * EMM_THREADS;
            const int qq = idx / KP, kk = idx - qq * KP;
            const bool v = idx < EMM_TQ * KP && (q0 + qq) < Q && kk < K;
            const size_t o = ((size_t)r * Q + q0 + qq) * K + kk;
            pf_mu[e] = v ? bf.mu_q[o] : ((idx < EMM_TQ * KP && kk == K && (q0 + qq) < Q) ? 1.0 : 0.0);
            pf_var[e] = v ? bf.var_q[o] : 0.0;
        }
        if (tid < EMM_TQ) pf_w = (q0 + tid < Q) ? bf.wq[(size_t)r * Q + q0 + tid] : 0.0;
    };
    if ((int)blockIdx.y < nitems) fetch(blockIdx.y);
    for (int it = blockIdx.y; it < nitems; it += gridDim.y) {
        const int rl = it / qtiles, r = dm.r0 + rl, q0 = (it - rl * qtiles) * EMM_TQ;
#pragma unroll
        for (int e = 0; e < NPF; ++e) {
            const int idx = tid + e * EMM_THREADS;
            if (idx < EMM_TQ * KP) {
                const int qq = idx / KP, kk = idx - qq * KP;
                muT[kk * LDQ + qq] = pf_mu[e];
                varT[kk * LDQ + qq] = pf_var[e];
            }
        }
        if (tid < EMM_TQ) ws[tid] = pf_w;
        __syncthreads();
        if (it + (int)gridDim.y < nitems) fetch(it + gridDim.y);
        // ---- GEMM A: h[qt][nl] = H[8 qt + g][8 (2 warp + nl) + 2 tg + e]
        {
            double h[2][2][2], sg[2][2][2];
#pragma unroll
            for (int nl = 0; nl < 2; ++nl) {
                const double2 d2 = *reinterpret_cast<const double2*>(dvec + 8 * (2 * warp + nl) + 2 * tg);
#pragma unroll
                for (int qt = 0; qt < 2; ++qt) {
                    h[qt][nl][0] = d2.x; h[qt][nl][1] = d2.y;
                    sg[qt][nl][0] = sg[qt][nl][1] = 0.0;
                }
            }
#pragma unroll
            for (int ks = 0; ks < KS4; ++ks) {
                double am[2], av[2], bc[2];
#pragma unroll
                for (int qt = 0; qt < 2; ++qt) {
                    am[qt] = muT[(4 * ks + tg) * LDQ + 8 * qt + g];
                    av[qt] = varT[(4 * ks + tg) * LDQ + 8 * qt + g];
                }
                double bc2[2];
#pragma unroll
                for (int nl = 0; nl < 2; ++nl) {
                    bc[nl] = CT[(4 * ks + tg) * TNS + 8 * (2 * warp + nl) + g];
                    bc2[nl] = CT2[(4 * ks + tg) * TNS + 8 * (2 * warp + nl) + g];
                }
#pragma unroll
                for (int qt = 0; qt < 2; ++qt)
#pragma unroll
                    for (int nl = 0; nl < 2; ++nl) {
                        dmma(h[qt][nl][0], h[qt][nl][1], am[qt], bc[nl]);
                        dmma(sg[qt][nl][0], sg[qt][nl][1], av[qt], bc2[nl]);
                    }
            }
#pragma unroll
            for (int qt = 0; qt < 2; ++qt) {
                const double w = ws[8 * qt + g];
#pragma unroll
                for (int nl = 0; nl < 2; ++nl) {
                    const int nn = 8 * (2 * warp + nl) + 2 * tg;
                    const double w0 = (n0 + nn < N) ? w : 0.0, w1 = (n0 + nn + 1 < N) ? w : 0.0;
                    // the ones column contributes muT[K] * CT[K] = 1 * 0 to h: nothing to undo
                    // (a table exp with four interleaved chains was measured here: 2.39 -> 2.56 ms on the 2000-trial shard --
                    //  the kernel sits at its 128-register budget for two resident CTAs and the chains spill)
                    const double e0 = w0 * exp(fma(0.5, sg[qt][nl][0], h[qt][nl][0]));
                    const double e1 = w1 * exp(fma(0.5, sg[qt][nl][1], h[qt][nl][1]));
                    t1 += e0 + e1;
                    *reinterpret_cast<double2*>(Gs + (8 * qt + g) * TNS + nn) = make_double2(-e0, -e1);
                }
            }
        }
        __syncthreads();
        // ---- GEMM B: 2 kinds x 2 point tiles x KT k-tiles, 32 k-steps over the 128 neurons
        if (need_lat) {
            for (int u = warp; u < 4 * KT; u += EMM_THREADS / 32) {
                const int kind = u & 1, qt = (u >> 1) & 1, kt = u >> 2;
                double c0 = 0.0, c1 = 0.0;
                const double* ga = Gs + (8 * qt + g) * TNS + tg;
                const double* cb = (kind ? CT2 : CT) + (8 * kt + g) * TNS + tg;
#pragma unroll 8
                for (int ks = 0; ks < EMM_TN / 4; ++ks) dmma(c0, c1, ga[4 * ks], cb[4 * ks]);
                const int q = q0 + 8 * qt + g, kk = 8 * kt + 2 * tg;
                if (q < Q) {
                    double* dst = kind ? bf.varbar_part : bf.mubar_part;
                    const double sc = kind ? 0.5 : 1.0;
                    if (kk < K) dst[part_off + ((size_t)r * K + kk) * Q + q] = sc * c0;
                    if (kk + 1 < K) dst[part_off + ((size_t)r * K + kk + 1) * Q + q] = sc * c1;
                }
            }
        }
        // ---- GEMM C: accumulate G^T Mu and G^T Var over the 16 points (4 k-steps)
        if (need_emb) {
#pragma unroll
            for (int ks = 0; ks < EMM_TQ / 4; ++ks) {
                double ga[2];
#pragma unroll
                for (int nl = 0; nl < 2; ++nl) ga[nl] = Gs[(4 * ks + tg) * TNS + 8 * (2 * warp + nl) + g];
#pragma unroll
                for (int kt = 0; kt < KT; ++kt) {
                    const double bm = muT[(8 * kt + g) * LDQ + 4 * ks + tg];
                    const double bv = varT[(8 * kt + g) * LDQ + 4 * ks + tg];
#pragma unroll
                    for (int nl = 0; nl < 2; ++nl) {
                        dmma(cm[nl][kt][0], cm[nl][kt][1], ga[nl], bm);
                        dmma(cv[nl][kt][0], cv[nl][kt][1], ga[nl], bv);
                    }
                }
            }
        }
        __syncthreads();
    }
    // ---- flush: dC[n][k] = cm + C o cv ; dd[n] = cm[.][K]
    if (need_emb) {
        double* gC = bf.shared + SVGPFA_SHARED_HDR;
        double* gd = gC + (size_t)N * K;
#pragma unroll
        for (int nl = 0; nl < 2; ++nl) {
            const int nn = 8 * (2 * warp + nl) + g, n = n0 + nn;
            if (n < N) {
#pragma unroll
                for (int kt = 0; kt < KT; ++kt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int kk = 8 * kt + 2 * tg + e;
                        if (kk < K) atomicAdd(gC + (size_t)n * K + kk, cm[nl][kt][e] + CT[kk * TNS + nn] * cv[nl][kt][e]);
                        else if (kk == K) atomicAdd(gd + n, cm[nl][kt][e]);
                    }
            }
        }
    }
    const double tot = block_sum(t1, red);
    if (tid == 0) {
        const int slot = (blockIdx.y * gridDim.x + blockIdx.x) % SVGPFA_TERM1_SLOTS;
        atomicAdd(bf.term1_part + slot, tot);
    }
}

// ------------------------------------------------------------------------------------------
// The same stage with the statistics CONCATENATED, for K <= 24 (round 2).  Mu and Var of a point enter every product
// side by side:  h = d + [Mu | Var] [C^T ; (C^T)^2 / 2]  is ONE product over 2K (one accumulator, ceil(2K/4) k-steps
// against two products over K + 1 padded to 8 each);  [mubar | varbar] = G [C | C^2/2]  and  G^T [Mu | Var]  have 2K output
// columns (K = 20: 5 column tiles against 3 + 3).  GEMM B is split over the NEURONS instead of over its output tiles:
// warp w multiplies the 16 x 16 block of G it has just produced (its own 16 neurons: no block barrier between A and B)
// into all 2 NC tiles -- 40 mma per warp at K = 20, all warps busy, where 12 tile-units of 32 mma over 8 warps took two
// rounds -- and the eight partial results meet in shared memory (upper four warps write, lower four add, then every
// thread sums four and writes the per-neuron-tile partials coalesced).  dd = column sums of G in registers (the ones
// column is gone).  mma per item and warp: 40 + 40 + 40 against 48 + 64 (critical path) + 48.
// ------------------------------------------------------------------------------------------
__host__ __device__ inline size_t emc_smem_doubles(int NC) {
    const int KP2 = 8 * NC;
    return (size_t)KP2 * EMM_TNS + (size_t)EMM_TQ * EMM_TNS + (size_t)KP2 * EMM_LDQ + EMM_TQ + EMM_TN + (size_t)8 * 2 * NC * 64;
}

template <int NC>      // NC = ceil(2 K / 8) column tiles of the concatenated statistics
__global__ void __launch_bounds__(EMM_THREADS, 2) quad_embed_cat_kernel(svgpfa_dims dm, svgpfa_buffers bf, uint32_t flags) {
    constexpr int KP2 = 8 * NC, KS = 2 * NC, TNS = EMM_TNS, LDQ = EMM_LDQ, NTB = 2 * NC;
    extern __shared__ __align__(16) double sm[];
    __shared__ double red[32];
    __shared__ double etab[64];
    svgpfa_load_exp_tab64(etab);
    const int K = dm.K, K2 = 2 * dm.K, N = dm.N, Q = dm.Q;
    double* CC = sm;                               // [KP2][TNS]  rows < K: C^T, rows K .. 2K-1: (C^T)^2 / 2, rest zero
    double* Gs = CC + (size_t)KP2 * TNS;           // [TQ][TNS]   G = -w exp(h); every warp reads only its own 16 columns
    double* stT = Gs + (size_t)EMM_TQ * TNS;       // [KP2][LDQ]  rows < K: Mu, rows K .. 2K-1: Var of the item's 16 points
    double* ws = stT + (size_t)KP2 * LDQ;          // [TQ]
    double* dvec = ws + EMM_TQ;                    // [TN]
    double* redB = dvec + EMM_TN;                  // [8 warps][NTB][32][2] partial GEMM-B tiles in fragment layout
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tg = lane & 3;
    const int tile = blockIdx.x, n0 = tile * EMM_TN;
    const bool need_emb = flags & SVGPFA_GRAD_EMBEDDING;
    const bool need_lat = flags & (SVGPFA_GRAD_POSTERIOR | SVGPFA_GRAD_KERNEL | SVGPFA_GRAD_INDLOCS);
    for (int idx = tid; idx < KP2 * EMM_TN; idx += EMM_THREADS) {
        const int c = idx / EMM_TN, nn = idx - c * EMM_TN;
        double v = 0.0;
        if (n0 + nn < N && c < K2) {
            const double cval = bf.C[(size_t)(n0 + nn) * K + (c < K ? c : c - K)];
            v = c < K ? cval : 0.5 * cval * cval;
        }
        CC[c * TNS + nn] = v;
    }
    if (tid < EMM_TN) dvec[tid] = (n0 + tid < N) ? bf.d[n0 + tid] : 0.0;
    // GEMM C accumulators: warp owns neuron tiles {2 warp, 2 warp + 1} x the NC column tiles of [Mu | Var]
    double cc[2][NC][2], dsum[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < NC; ++b) cc[a][b][0] = cc[a][b][1] = 0.0;
    double t1 = 0.0;
    const int qtiles = (Q + EMM_TQ - 1) / EMM_TQ;
    const int nitems = (dm.rn ? dm.rn : dm.R) * qtiles;
    const size_t part_off = (size_t)tile * dm.R * K * Q;
    __syncthreads();
    constexpr int NPF = (EMM_TQ * KP2 + EMM_THREADS - 1) / EMM_THREADS;
    double pf[NPF], pf_w = 0.0;
    auto fetch = [&](int it) {                    // the NEXT item's statistics travel while the current products run
        const int rl = it / qtiles, r = dm.r0 + rl, q0 = (it - rl * qtiles) * EMM_TQ;
#pragma unroll
        for (int e = 0; e < NPF; ++e) {
            const int idx = tid + e * EMM_THREADS;
            const int qq = idx / KP2, c = idx - qq * KP2;
            const bool v = idx < EMM_TQ * KP2 && (q0 + qq) < Q && c < K2;
            const size_t o = ((size_t)r * Q + q0 + qq) * K;
            pf[e] = v ? (c < K ? bf.mu_q[o + c] : bf.var_q[o + c - K]) : 0.0;
        }
        if (tid < EMM_TQ) pf_w = (q0 + tid < Q) ? bf.wq[(size_t)r * Q + q0 + tid] : 0.0;
    };
    if ((int)blockIdx.y < nitems) fetch(blockIdx.y);
    for (int it = blockIdx.y; it < nitems; it += gridDim.y) {
        const int rl = it / qtiles, r = dm.r0 + rl, q0 = (it - rl * qtiles) * EMM_TQ;
#pragma unroll
        for (int e = 0; e < NPF; ++e) {
            const int idx = tid + e * EMM_THREADS;
            if (idx < EMM_TQ * KP2) {
                const int qq = idx / KP2, c = idx - qq * KP2;
                stT[c * LDQ + qq] = pf[e];
            }
        }
        if (tid < EMM_TQ) ws[tid] = pf_w;
        __syncthreads();
        if (it + (int)gridDim.y < nitems) fetch(it + gridDim.y);
        // ---- GEMM A: h[qt][nl] = H[8 qt + g][8 (2 warp + nl) + 2 tg + e] = d + [Mu | Var] CC
        {
            double h[2][2][2];
#pragma unroll
            for (int nl = 0; nl < 2; ++nl) {
                const double2 d2 = *reinterpret_cast<const double2*>(dvec + 8 * (2 * warp + nl) + 2 * tg);
#pragma unroll
                for (int qt = 0; qt < 2; ++qt) { h[qt][nl][0] = d2.x; h[qt][nl][1] = d2.y; }
            }
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                double a[2], b[2];
#pragma unroll
                for (int qt = 0; qt < 2; ++qt) a[qt] = stT[(4 * ks + tg) * LDQ + 8 * qt + g];
#pragma unroll
                for (int nl = 0; nl < 2; ++nl) b[nl] = CC[(4 * ks + tg) * TNS + 8 * (2 * warp + nl) + g];
#pragma unroll
                for (int qt = 0; qt < 2; ++qt)
#pragma unroll
                    for (int nl = 0; nl < 2; ++nl) dmma(h[qt][nl][0], h[qt][nl][1], a[qt], b[nl]);
            }
#pragma unroll
            for (int qt = 0; qt < 2; ++qt) {
                const double w = ws[8 * qt + g];
                // table exp, four interleaved evaluations, where the registers allow it (NC <= 4: config #3 0.714 -> 0.694 ms
                // on the 2000-trial shard; NC = 5 spills and loses, 2.17 -> 2.22: libdevice there)
                double hx[4] = {h[qt][0][0], h[qt][0][1], h[qt][1][0], h[qt][1][1]}, ex[4];
                if (NC <= 4) {
                    svgpfa_exp_neg64_n<4, true>(hx, etab, ex);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) ex[e] = exp(hx[e]);
                }
#pragma unroll
                for (int nl = 0; nl < 2; ++nl) {
                    const int nn = 8 * (2 * warp + nl) + 2 * tg;
                    const double w0 = (n0 + nn < N) ? w : 0.0, w1 = (n0 + nn + 1 < N) ? w : 0.0;
                    const double e0 = w0 * ex[2 * nl];
                    const double e1 = w1 * ex[2 * nl + 1];
                    t1 += e0 + e1;
                    dsum[nl][0] -= e0;
                    dsum[nl][1] -= e1;
                    *reinterpret_cast<double2*>(Gs + (8 * qt + g) * TNS + nn) = make_double2(-e0, -e1);
                }
            }
        }
        __syncwarp();                             // this warp's 16 columns of G are complete; nobody else reads them
        // ---- GEMM B, this warp's 16 neurons: pb[qt][ct] = sum_n G[8 qt + g][n] CC[8 ct + 2 tg + e][n]
        if (need_lat) {
            double pb[2][NC][2];
#pragma unroll
            for (int qt = 0; qt < 2; ++qt)
#pragma unroll
                for (int ct = 0; ct < NC; ++ct) pb[qt][ct][0] = pb[qt][ct][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const int n = 16 * warp + 4 * ks + tg;
                double ga[2];
#pragma unroll
                for (int qt = 0; qt < 2; ++qt) ga[qt] = Gs[(8 * qt + g) * TNS + n];
#pragma unroll
                for (int ct = 0; ct < NC; ++ct) {
                    const double b = CC[(8 * ct + g) * TNS + n];
#pragma unroll
                    for (int qt = 0; qt < 2; ++qt) dmma(pb[qt][ct][0], pb[qt][ct][1], ga[qt], b);
                }
            }
            // the eight partial results meet in shared memory, in fragment layout (one conflict-free 16-byte store per tile)
            double* mine = redB + (size_t)warp * NTB * 64 + 2 * lane;
#pragma unroll
            for (int qt = 0; qt < 2; ++qt)
#pragma unroll
                for (int ct = 0; ct < NC; ++ct)
                    *reinterpret_cast<double2*>(mine + (qt * NC + ct) * 64) = make_double2(pb[qt][ct][0], pb[qt][ct][1]);
            // ---- GEMM C in between (independent of the exchange): accumulate G^T [Mu | Var] over the 16 points
            if (need_emb) {
#pragma unroll
                for (int ks = 0; ks < EMM_TQ / 4; ++ks) {
                    double ga[2];
#pragma unroll
                    for (int nl = 0; nl < 2; ++nl) ga[nl] = Gs[(4 * ks + tg) * TNS + 8 * (2 * warp + nl) + g];
#pragma unroll
                    for (int ct = 0; ct < NC; ++ct) {
                        const double b = stT[(8 * ct + g) * LDQ + 4 * ks + tg];
#pragma unroll
                        for (int nl = 0; nl < 2; ++nl) dmma(cc[nl][ct][0], cc[nl][ct][1], ga[nl], b);
                    }
                }
            }
            __syncthreads();                      // everyone is done with stT; all partials are in place
            // 8 -> 1 and out: thread <-> (tile, fragment lane); rows of 8 consecutive points per (column, g-run)
            for (int u = tid; u < NTB * 32; u += EMM_THREADS) {
                const double2* p = reinterpret_cast<const double2*>(redB + 2 * u);
                double2 s = p[0];
#pragma unroll
                for (int w = 1; w < 8; ++w) {
                    const double2 x = p[w * NTB * 32];
                    s.x += x.x;
                    s.y += x.y;
                }
                const int tl = u >> 5, ln = u & 31, qt = tl / NC, ct = tl - qt * NC;
                const int q = q0 + 8 * qt + (ln >> 2), c = 8 * ct + 2 * (ln & 3);
                if (q < Q) {
                    if (c < K) bf.mubar_part[part_off + ((size_t)r * K + c) * Q + q] = s.x;
                    else if (c < K2) bf.varbar_part[part_off + ((size_t)r * K + c - K) * Q + q] = s.x;
                    if (c + 1 < K) bf.mubar_part[part_off + ((size_t)r * K + c + 1) * Q + q] = s.y;
                    else if (c + 1 < K2) bf.varbar_part[part_off + ((size_t)r * K + c + 1 - K) * Q + q] = s.y;
                }
            }
        } else {
            if (need_emb) {
#pragma unroll
                for (int ks = 0; ks < EMM_TQ / 4; ++ks) {
                    double ga[2];
#pragma unroll
                    for (int nl = 0; nl < 2; ++nl) ga[nl] = Gs[(4 * ks + tg) * TNS + 8 * (2 * warp + nl) + g];
#pragma unroll
                    for (int ct = 0; ct < NC; ++ct) {
                        const double b = stT[(8 * ct + g) * LDQ + 4 * ks + tg];
#pragma unroll
                        for (int nl = 0; nl < 2; ++nl) dmma(cc[nl][ct][0], cc[nl][ct][1], ga[nl], b);
                    }
                }
            }
            __syncthreads();                      // everyone is done with stT
        }
    }
    // ---- flush: dC[n][k] = (G^T Mu)[n][k] + C[n][k] (G^T Var)[n][k];  dd[n] = sum_q G[q][n]
    if (need_emb) {
        double* gC = bf.shared + SVGPFA_SHARED_HDR;
        double* gd = gC + (size_t)N * K;
#pragma unroll
        for (int nl = 0; nl < 2; ++nl) {
            const int nn = 8 * (2 * warp + nl) + g, n = n0 + nn;
            if (n < N) {
#pragma unroll
                for (int ct = 0; ct < NC; ++ct)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int c = 8 * ct + 2 * tg + e;
                        if (c < K) atomicAdd(gC + (size_t)n * K + c, cc[nl][ct][e]);
                        else if (c < K2) atomicAdd(gC + (size_t)n * K + c - K, CC[(c - K) * TNS + nn] * cc[nl][ct][e]);
                    }
            }
            // dsum[nl][e]: this lane's points of neuron 8 (2 warp + nl) + 2 tg + e; the other points sit in the lanes g' != g
            double s0 = dsum[nl][0], s1 = dsum[nl][1];
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                s0 += __shfl_xor_sync(0xffffffffu, s0, o);
                s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            }
            const int nd = n0 + 8 * (2 * warp + nl) + 2 * tg;
            if (g == 0) {
                if (nd < N) atomicAdd(gd + nd, s0);
                if (nd + 1 < N) atomicAdd(gd + nd + 1, s1);
            }
        }
    }
    const double tot = block_sum(t1, red);
    if (tid == 0) {
        const int slot = (blockIdx.y * gridDim.x + blockIdx.x) % SVGPFA_TERM1_SLOTS;
        atomicAdd(bf.term1_part + slot, tot);
    }
}

template <int NC>
void launch_emc(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, cudaStream_t st) {
    const size_t smem = sizeof(double) * emc_smem_doubles(NC);
    static std::atomic<int> occ_cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    int occ = occ_cache[dev & 63].load(std::memory_order_relaxed);
    if (occ <= 0) {
        cudaFuncSetAttribute(quad_embed_cat_kernel<NC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(quad_embed_cat_kernel<NC>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, quad_embed_cat_kernel<NC>, EMM_THREADS, smem);
        if (occ < 1) occ = 1;
        occ_cache[dev & 63].store(occ, std::memory_order_relaxed);
    }
    const int ntiles = (dims->N + EMM_TN - 1) / EMM_TN;
    const int qtiles = (dims->Q + EMM_TQ - 1) / EMM_TQ;
    const long nitems = (long)svgpfa_ntrials(dims) * qtiles;
    long workers = (long)svgpfa_sm_count() * occ / ntiles;
    if (workers < 1) workers = 1;
    if (workers > nitems) workers = nitems;
    quad_embed_cat_kernel<NC><<<dim3(ntiles, (unsigned)workers), EMM_THREADS, smem, st>>>(*dims, *buf, flags);
}

template <int KT>
void launch_emm(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, cudaStream_t st) {
    // shared memory depends on KP = 8 KT only, so the attribute and the occupancy are fixed per instantiation
    const size_t smem = sizeof(double) * emm_smem_doubles(8 * KT - 1);
    static std::atomic<int> occ_cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    int occ = occ_cache[dev & 63].load(std::memory_order_relaxed);
    if (occ <= 0) {
        cudaFuncSetAttribute(quad_embed_mma_kernel<KT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(quad_embed_mma_kernel<KT>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, quad_embed_mma_kernel<KT>, EMM_THREADS, smem);
        if (occ < 1) occ = 1;
        occ_cache[dev & 63].store(occ, std::memory_order_relaxed);
    }
    const int ntiles = (dims->N + EMM_TN - 1) / EMM_TN;
    const int qtiles = (dims->Q + EMM_TQ - 1) / EMM_TQ;
    const long nitems = (long)svgpfa_ntrials(dims) * qtiles;
    const int nsm = svgpfa_sm_count();
    long workers = (long)nsm * occ / ntiles;
    if (workers < 1) workers = 1;
    if (workers > nitems) workers = nitems;
    quad_embed_mma_kernel<KT><<<dim3(ntiles, (unsigned)workers), EMM_THREADS, smem, st>>>(*dims, *buf, flags);
}

}  // namespace

// Returns false when K is outside this path (K + 1 > 40 latents); the caller then uses the CUDA-core kernel.
bool svgpfa_try_quad_embed_mma(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, cudaStream_t st) {
    switch ((2 * dims->K + 7) / 8) {                  // K <= 24: the concatenated-statistics kernel
        case 1: launch_emc<1>(dims, buf, flags, st); return true;
        case 2: launch_emc<2>(dims, buf, flags, st); return true;
        case 3: launch_emc<3>(dims, buf, flags, st); return true;
        case 4: launch_emc<4>(dims, buf, flags, st); return true;
        case 5: launch_emc<5>(dims, buf, flags, st); return true;
        case 6: launch_emc<6>(dims, buf, flags, st); return true;
        default: break;
    }
    const int KT = emm_kp(dims->K) / 8;
    switch (KT) {
        case 1: launch_emm<1>(dims, buf, flags, st); break;
        case 2: launch_emm<2>(dims, buf, flags, st); break;
        case 3: launch_emm<3>(dims, buf, flags, st); break;
        case 4: launch_emm<4>(dims, buf, flags, st); break;
        case 5: launch_emm<5>(dims, buf, flags, st); break;
        default: return false;
    }
    return true;
}

// Returns false when the shape is outside this path (M > 32); the caller then uses the CUDA-core kernels.
bool svgpfa_try_quad_latent_mma(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, bool bwd,
                                cudaStream_t st) {
    const int M = dims->Mmax;
    if (M > 32) return false;
    const int MT = (M + 7) / 8;
    if (bwd) {
        switch (MT) {
            case 1: launch_qm<1, true>(dims, buf, flags, st); break;
            case 2: launch_qm<2, true>(dims, buf, flags, st); break;
            case 3: launch_qm<3, true>(dims, buf, flags, st); break;
            default: launch_qm<4, true>(dims, buf, flags, st); break;
        }
    } else {
        switch (MT) {
            case 1: launch_qm<1, false>(dims, buf, flags, st); break;
            case 2: launch_qm<2, false>(dims, buf, flags, st); break;
            case 3: launch_qm<3, false>(dims, buf, flags, st); break;
            default: launch_qm<4, false>(dims, buf, flags, st); break;
        }
    }
    return true;
}

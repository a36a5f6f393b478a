// Latent posterior at the quadrature points and its adjoint for 32 < M <= 64 on the FP64 tensor path (round 2).
// Same arithmetic as quad_mma.cu (reference: stats/kernelsMatricesStore.py:186-195,
// stats/svPosteriorOnLatents.py:185-216); what changes is who owns what.
//
// For M <= 32 a warp owns a tile of points and carries all M rows of V, U, W (and its own partial of the M x M matrix
// A = sum_q varbar_q v_q v_q^T) in registers.  At M = 64 that is 4x the accumulators: they do not fit.  Here the CTA
// (8 warps) walks the points in passes of 32 TOGETHER and the ROWS are dealt to the warps:
//   * products with an M x 32 result (V = Li K, U = X^T V, W = G V, Kv = Li^T W): warp w owns the row tiles
//     {p, MT-1-p}, p = w / 2 -- a pair has the same number of k-steps in every triangular product -- and the column
//     half w & 1: 8 accumulator doubles.  The operand tile lives in shared memory once per CTA; every product reads it
//     completely before its result is written back over it (block barriers on either side);
//   * A: the 36 lower tiles are dealt round-robin to the warps and each stays in ONE warp's registers for the whole
//     kernel -- no cross-warp sum, the tile is stored straight from the fragments at the end;
//   * element-wise phases (kernel values, the adjoints of the kernel evaluations): 256 threads over
//     (inducing point, quarter of the pass's points) or (point, block of 8 inducing points), four interleaved evaluation
//     chains (kappa_vals_n).
// The adjoint REQUIRES the V cache (buffers.v_q, written by the forward kernel): it never evaluates K for V, and
// abar = L sum_q mubar_q v_q (see quad_mma.cu).  Without the cache the caller falls back to the CUDA-core kernels of
// quad.cu, which these kernels replace as the M > 32 path (26.4 / 99.6 ms -> see profiles/README.md, R = 4000, M = 64).
#include <type_traits>

#include "common.cuh"

namespace {

constexpr int QB_WARPS = 8, QB_THREADS = 32 * QB_WARPS, QB_LDT = 36;

__device__ __forceinline__ void dmma_b(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

__host__ __device__ inline size_t qb_smem_doubles(int MP) {
    return (size_t)2 * MP * (MP + 4) + (size_t)MP * QB_LDT + 3 * MP + 3 * 32 + 3 * QB_WARPS * 32;
}

// row tiles of this warp in the M x 32 products: {p, MT-1-p} (one tile when they coincide), none for idle warps
template <int MT>
__device__ __forceinline__ int qb_rows(int warp, int (&rt)[2]) {
    const int p = warp >> 1;
    if (p >= (MT + 1) / 2) return 0;
    rt[0] = p;
    rt[1] = MT - 1 - p;
    return rt[0] == rt[1] ? 1 : 2;
}

template <int MT, bool BWD, bool VC>
__global__ void __launch_bounds__(QB_THREADS, 2) quad_latent_big_kernel(svgpfa_dims dm, svgpfa_buffers bf, uint32_t flags) {
    constexpr int MP = 8 * MT, KS = 2 * MT, LD = MP + 4, LDT = QB_LDT, NTA = MT * (MT + 1) / 2;
    constexpr int NOWN = (NTA + QB_WARPS - 1) / QB_WARPS;             // tiles of A per warp
    extern __shared__ __align__(16) double sm[];
    __shared__ double red[32];
    __shared__ double etab[64];
    __shared__ double2 sctab[SVGPFA_SC_ENTRIES];
    svgpfa_load_exp_tab64(etab);
    const int r = dm.r0 + blockIdx.x, k = blockIdx.y;
    const svgpfa_latent_desc ds = bf.desc[k];
    const int M = ds.M, Q = dm.Q;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tg = lane & 3;
    const bool need_kz = BWD && (flags & (SVGPFA_GRAD_KERNEL | SVGPFA_GRAD_INDLOCS));
    double* Lis = sm;                                   // [MP][LD] Li  (end of the adjoint: L)
    double* Xs = Lis + MP * LD;                         // [MP][LD] X;  adjoint with need_kz: G = X X^T - I
    double* T = Xs + MP * LD;                           // [MP][LDT] the pass's tile: K, V, (W, Li^T W)
    double* zs = T + MP * LDT;                          // [MP]
    double* al = zs + MP;                               // [MP] alpha (forward) / vm at the end of the adjoint
    double* vmx = al + MP;                              // [MP] scratch vector
    double* tt = vmx + MP;                              // [32] nodes of the pass
    double* mbs = tt + 32;                              // [32] mubar
    double* vbs = mbs + 32;                             // [32] varbar
    double* part = vbs + 32;                            // [3][8 warps][32] partial column sums (mu, |v|^2, |u|^2)
    const KConst kc = make_kconst(ds, bf.theta, bf.kscale, k);
    if (kc.type == SVGPFA_KERNEL_PERIODIC) svgpfa_load_sincos_tab<1>(sctab);
    const size_t mo = (size_t)r * dm.MM + ds.mmoff, vo = (size_t)r * dm.KM + ds.moff;
    for (int idx = tid; idx < MP * MP; idx += QB_THREADS) {
        const int i = idx / MP, j = idx - i * MP;
        const bool in = i < M && j < M;
        Lis[i * LD + j] = in ? bf.Li[mo + (size_t)i * M + j] : 0.0;
        Xs[i * LD + j] = in ? bf.X[mo + (size_t)i * M + j] : 0.0;
    }
    const double* zg = bf.Z + (size_t)dm.R * ds.moff + (size_t)r * M;
    for (int i = tid; i < MP; i += QB_THREADS) {
        zs[i] = i < M ? zg[i] : 0.0;
        al[i] = i < M ? bf.alpha[vo + i] : 0.0;
    }
    __syncthreads();
    int rts[2];
    const int nrt = qb_rows<MT>(warp, rts);
    const int ch = warp & 1;                            // column half: point tiles 2 ch, 2 ch + 1 of the pass
    if (need_kz) {
        // G = X X^T - I over X: warp w computes row tiles w, w + 8, ... (all columns) into registers, then writes
        double gacc[MT][2];
        for (int it = warp; it < MT; it += QB_WARPS) {
#pragma unroll
            for (int jt = 0; jt < MT; ++jt) gacc[jt][0] = gacc[jt][1] = 0.0;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks) {
                const double a = Xs[(8 * it + g) * LD + 4 * ks + tg];
#pragma unroll
                for (int jt = 0; jt < MT; ++jt) dmma_b(gacc[jt][0], gacc[jt][1], a, Xs[(8 * jt + g) * LD + 4 * ks + tg]);
            }
        }
        __syncthreads();
        for (int it = warp; it < MT; it += QB_WARPS) {
#pragma unroll
            for (int jt = 0; jt < MT; ++jt) {
                const int i = 8 * it + g, j = 8 * jt + 2 * tg;
                *reinterpret_cast<double2*>(Xs + i * LD + j) =
                    make_double2(gacc[jt][0] - (i == j ? 1.0 : 0.0), gacc[jt][1] - (i == j + 1 ? 1.0 : 0.0));
            }
        }
        __syncthreads();
    }
    // this warp's tiles of A (adjoint): tile index t = w, w + 8, ... in the order (0,0), (1,0), (1,1), (2,0), ...
    double accA[BWD ? NOWN : 1][2];
    int a_it[BWD ? NOWN : 1], a_jt[BWD ? NOWN : 1];
#pragma unroll
    for (int o = 0; o < (BWD ? NOWN : 1); ++o) {
        accA[o][0] = accA[o][1] = 0.0;
        int t = warp + o * QB_WARPS, it = 0;
        if (t >= NTA) t = -1;
        if (t >= 0) while (t >= it + 1) { t -= it + 1; ++it; }
        a_it[o] = t >= 0 ? it : -1;
        a_jt[o] = t;
    }
    // element-wise mapping of the adjoint: thread <-> (inducing point je, quarter of the pass's points)
    const int je = tid & 63, qpart = tid >> 6;
    double vm_own = 0.0, dz_raw = 0.0, th0_raw = 0.0, th1_raw = 0.0;
    const double zj_own = zs[je < MP ? je : 0], aj_own = al[je < MP ? je : 0];
    const size_t part_stride = (size_t)dm.R * dm.K * Q;
    const double* vg = VC ? bf.v_q + ((size_t)r * dm.KM + ds.moff) * Q : nullptr;

    // one M x 32 product into registers: acc[s][c] = rows rts[s], columns 8 (2 ch + c) + 2 tg + {0, 1}
    //   KIND 0: Li (lower) x T      1: X^T (upper) x T      2: G (full) x T      3: Li^T (upper) x T
    auto product = [&](auto kind_tag, double (&acc)[2][2][2]) {
        constexpr int KIND = decltype(kind_tag)::value;
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int c = 0; c < 2; ++c) acc[s][c][0] = acc[s][c][1] = 0.0;
        for (int s = 0; s < nrt; ++s) {
            const int rt = rts[s];
            const int lo = (KIND == 1 || KIND == 3) ? 2 * rt : 0, hi = KIND == 0 ? 2 * rt + 1 : KS - 1;
#pragma unroll 4
            for (int ks = lo; ks <= hi; ++ks) {
                double a;
                if (KIND == 0) a = Lis[(8 * rt + g) * LD + 4 * ks + tg];
                else if (KIND == 1) a = Xs[(4 * ks + tg) * LD + 8 * rt + g];
                else if (KIND == 2) a = Xs[(8 * rt + g) * LD + 4 * ks + tg];
                else a = Lis[(4 * ks + tg) * LD + 8 * rt + g];
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    const double b = T[(4 * ks + tg) * LDT + 8 * (2 * ch + c) + g];
                    if (s == 0) dmma_b(acc[0][c][0], acc[0][c][1], a, b);
                    else dmma_b(acc[1][c][0], acc[1][c][1], a, b);
                }
            }
        }
    };
    auto store_tile = [&](const double (&acc)[2][2][2]) {
        for (int s = 0; s < nrt; ++s)
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                double2 v2 = s == 0 ? make_double2(acc[0][c][0], acc[0][c][1]) : make_double2(acc[1][c][0], acc[1][c][1]);
                *reinterpret_cast<double2*>(T + (8 * rts[s] + g) * LDT + 8 * (2 * ch + c) + 2 * tg) = v2;
            }
    };
    // column sums of the squares of this warp's rows, for its 16 columns -> part[which][warp][col] (zero elsewhere)
    auto colsq = [&](const double (&acc)[2][2][2], int which) {
        part[(which * QB_WARPS + warp) * 32 + lane] = 0.0;
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            double d0 = 0.0, d1 = 0.0;
            for (int s = 0; s < nrt; ++s) {
                const double x0 = s == 0 ? acc[0][c][0] : acc[1][c][0], x1 = s == 0 ? acc[0][c][1] : acc[1][c][1];
                d0 = fma(x0, x0, d0);
                d1 = fma(x1, x1, d1);
            }
#pragma unroll
            for (int o = 4; o < 32; o <<= 1) {
                d0 += __shfl_xor_sync(0xffffffffu, d0, o);
                d1 += __shfl_xor_sync(0xffffffffu, d1, o);
            }
            if (g == 0) {
                part[(which * QB_WARPS + warp) * 32 + 8 * (2 * ch + c) + 2 * tg] = d0;
                part[(which * QB_WARPS + warp) * 32 + 8 * (2 * ch + c) + 2 * tg + 1] = d1;
            }
        }
    };

    const int npass = (Q + 31) / 32;
    for (int ps = 0; ps < npass; ++ps) {
        const int qbase = 32 * ps;
        double acc[2][2][2];
        if (!BWD) {
            // ---- K[j][q] -> T, thread <-> (point lane, inducing points 8 warp .. 8 warp + 7); mu partials
            const int q = qbase + lane;
            const bool valid = q < Q;
            const double t_lane = valid ? bf.tq[(size_t)r * Q + q] : 0.0;
            double mu = 0.0;
            if (warp < MT) {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int j0 = 8 * warp + 4 * h;
                    double dl[4], kv[4], qq[4], s2x[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) dl[e] = t_lane - zs[j0 + e];
                    kappa_vals_n<4>(kc, dl, etab, sctab, kv, qq, s2x);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const double v = (valid && j0 + e < M) ? kv[e] : 0.0;
                        T[(j0 + e) * LDT + lane] = v;
                        mu = fma(v, al[j0 + e], mu);
                    }
                }
            }
            part[warp * 32 + lane] = mu;
            __syncthreads();
            if (warp == 0 && valid) {
                double s = 0.0;
#pragma unroll
                for (int w = 0; w < QB_WARPS; ++w) s += part[w * 32 + lane];
                bf.mu_q[((size_t)r * Q + q) * dm.K + k] = s;
            }
            // ---- V = Li K
            product(std::integral_constant<int, 0>{}, acc);
            if (VC) {
                double* vgw = bf.v_q + ((size_t)r * dm.KM + ds.moff) * Q;
                for (int s = 0; s < nrt; ++s)
#pragma unroll
                    for (int c = 0; c < 2; ++c) {
                        const int row = 8 * rts[s] + g, col = qbase + 8 * (2 * ch + c) + 2 * tg;
                        const double2 v2 = s == 0 ? make_double2(acc[0][c][0], acc[0][c][1]) : make_double2(acc[1][c][0], acc[1][c][1]);
                        if (row < M && col < Q) *reinterpret_cast<double2*>(vgw + (size_t)row * Q + col) = v2;
                    }
            }
            colsq(acc, 1);
            __syncthreads();                                         // every warp is done reading K
            store_tile(acc);
            __syncthreads();
            // ---- U = X^T V, var = s2 - |v|^2 + |u|^2
            product(std::integral_constant<int, 1>{}, acc);
            colsq(acc, 2);
            __syncthreads();                                         // V consumed, partial sums in place
            if (warp == 0 && valid) {
                double s = kc.s2;
#pragma unroll
                for (int w = 0; w < QB_WARPS; ++w) s += part[(2 * QB_WARPS + w) * 32 + lane] - part[(QB_WARPS + w) * 32 + lane];
                bf.var_q[((size_t)r * Q + q) * dm.K + k] = s;
            }
            __syncthreads();                                         // part is rewritten by the next pass
        } else {
            // ---- the tile receives V from the cache (zero fill for rows >= M and points >= Q)
            {
                const unsigned tile_s = (unsigned)__cvta_generic_to_shared(T);
                for (int c = tid; c < MP * 16; c += QB_THREADS) {
                    const int j = c >> 4, cl = c & 15, col = qbase + 2 * cl;
                    const bool ok = j < M && col < Q;
                    const double* src = vg + (ok ? (size_t)j * Q + col : 0);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(tile_s + (unsigned)(j * LDT + 2 * cl) * 8u), "l"(src),
                                 "r"(ok ? 16 : 0) : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            if (tid < 32) {
                const int q = qbase + tid;
                double mbar = 0.0, vbar = 0.0, tq = 0.0;
                if (q < Q) {
                    const size_t o = ((size_t)r * dm.K + k) * Q + q;
                    for (int p = 0; p < dm.n_ntiles; ++p) {
                        mbar += bf.mubar_part[p * part_stride + o];
                        vbar += bf.varbar_part[p * part_stride + o];
                    }
                    tq = bf.tq[(size_t)r * Q + q];
                }
                mbs[tid] = mbar;
                vbs[tid] = vbar;
                tt[tid] = tq;
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
            __syncthreads();
            // ---- vm_j += sum_q mubar_q V[j][q], this thread's quarter of the points
            if (je < MP) {
                double s0 = 0.0, s1 = 0.0;
#pragma unroll
                for (int c = 0; c < 8; c += 2) {
                    const int q0 = 8 * qpart + ((c + je) & 7), q1 = 8 * qpart + ((c + 1 + je) & 7);
                    s0 = fma(mbs[q0], T[je * LDT + q0], s0);
                    s1 = fma(mbs[q1], T[je * LDT + q1], s1);
                }
                vm_own += s0 + s1;
            }
            // ---- A += V diag(varbar) V^T, this warp's tiles
#pragma unroll
            for (int o = 0; o < NOWN; ++o) {
                if (a_it[o] >= 0) {
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {
                        const double a = T[(8 * a_it[o] + g) * LDT + 4 * ks + tg] * vbs[4 * ks + tg];
                        dmma_b(accA[o][0], accA[o][1], a, T[(8 * a_jt[o] + g) * LDT + 4 * ks + tg]);
                    }
                }
            }
            if (need_kz) {
                // ---- W = G V -> T, Kv = Li^T W -> T
                product(std::integral_constant<int, 2>{}, acc);
                __syncthreads();                                     // V consumed by every warp (vm, A, W)
                store_tile(acc);
                __syncthreads();
                product(std::integral_constant<int, 3>{}, acc);
                __syncthreads();
                store_tile(acc);
                __syncthreads();
                // ---- kbar = 2 varbar Kv + mubar alpha and its products with dkappa (raw moments, see quad_mma.cu)
                if (je < M) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        double dl[4], kv[4], qq[4], s2x[4];
                        int qi[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            qi[e] = 8 * qpart + ((4 * h + e + je) & 7);
                            dl[e] = tt[qi[e]] - zj_own;
                        }
                        kappa_vals_n<4>(kc, dl, etab, sctab, kv, qq, s2x);
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const double kbar = fma(2.0 * vbs[qi[e]], T[je * LDT + qi[e]], mbs[qi[e]] * aj_own);
                            const double hh = (qbase + qi[e] < Q) ? kbar * kv[e] : 0.0;
                            if (kc.type == SVGPFA_KERNEL_EXPQUAD) {
                                dz_raw = fma(hh, dl[e], dz_raw);
                                th0_raw = fma(hh, qq[e], th0_raw);
                            } else {
                                const double hs = hh * s2x[e];
                                dz_raw += hs;
                                th0_raw = fma(hh, qq[e], th0_raw);
                                th1_raw = fma(hs, dl[e], th1_raw);
                            }
                        }
                    }
                }
            }
            __syncthreads();                                         // the tile and mbs / vbs / tt are rewritten by the next pass
        }
    }
    if (!BWD) return;
    // ---- A_q: every tile straight from its owner's fragments (lower triangle incl. diagonal)
#pragma unroll
    for (int o = 0; o < NOWN; ++o) {
        if (a_it[o] >= 0) {
            const int i = 8 * a_it[o] + g, j = 8 * a_jt[o] + 2 * tg;
            if (i < M) {
                if (j <= i) bf.A_q[mo + (size_t)i * M + j] = accA[o][0];
                if (j + 1 <= i) bf.A_q[mo + (size_t)i * M + j + 1] = accA[o][1];
            }
        }
    }
    // ---- abar = L vm, dz, dtheta: the four point-quarters of an inducing point meet in shared memory
    const double dz_own = -kc.dd * dz_raw, th0 = kc.dl * th0_raw, th1 = kc.dp * th1_raw;
    if (je < MP) {
        T[qpart * MP + je] = vm_own;
        T[(4 + qpart) * MP + je] = dz_own;
    }
    for (int idx = tid; idx < MP * MP; idx += QB_THREADS) {              // L over Li (all products are done)
        const int i = idx / MP, j = idx - i * MP;
        Lis[i * LD + j] = (i < M && j <= i) ? bf.L[mo + (size_t)i * M + j] : 0.0;
    }
    __syncthreads();
    if (tid < MP) {
        vmx[tid] = (T[tid] + T[MP + tid]) + (T[2 * MP + tid] + T[3 * MP + tid]);
        const double sz = (T[4 * MP + tid] + T[5 * MP + tid]) + (T[6 * MP + tid] + T[7 * MP + tid]);
        if (need_kz && tid < M) atomicAdd(bf.dz_acc + vo + tid, sz);
    }
    __syncthreads();
    if (tid < M) {
        double s0 = 0.0, s1 = 0.0;
#pragma unroll 4
        for (int p = 0; p < MP; p += 2) {
            s0 = fma(Lis[tid * LD + p], vmx[p], s0);
            s1 = fma(Lis[tid * LD + p + 1], vmx[p + 1], s1);
        }
        bf.abar_q[vo + tid] = s0 + s1;
    }
    if (need_kz && (flags & SVGPFA_GRAD_KERNEL)) {
        const double s0 = block_sum(th0, red);
        const double s1 = block_sum(th1, red);
        if (tid == 0) {
            double* dth = bf.dth_part + (size_t)r * dm.TH + ds.thoff;
            atomicAdd(dth, s0);
            if (ds.nth > 1) atomicAdd(dth + 1, s1);
        }
    }
}

template <int MT, bool BWD, bool VC>
void launch_qb(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, cudaStream_t st) {
    const size_t smem = sizeof(double) * qb_smem_doubles(8 * MT);
    SVGPFA_ENSURE_SMEM(smem, quad_latent_big_kernel<MT, BWD, VC>);
    quad_latent_big_kernel<MT, BWD, VC><<<dim3(svgpfa_ntrials(dims), dims->K), QB_THREADS, smem, st>>>(*dims, *buf, flags);
}

template <int MT>
bool launch_qb_mt(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, bool bwd, cudaStream_t st) {
    const bool vc = buf->v_q && (dims->Q & 1) == 0;
    if (bwd) {
        if (!vc) return false;                          // the adjoint needs the V cache: CUDA-core fallback otherwise
        launch_qb<MT, true, true>(dims, buf, flags, st);
    } else if (vc && !(flags & SVGPFA_REUSE_VQ)) {
        launch_qb<MT, false, true>(dims, buf, flags, st);
    } else {
        launch_qb<MT, false, false>(dims, buf, flags, st);
    }
    return true;
}

}  // namespace

// 32 < Mmax <= 64.  Returns false when this path does not apply (the caller then uses the CUDA-core kernels).
bool svgpfa_try_quad_latent_big(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, bool bwd, cudaStream_t st) {
    switch ((dims->Mmax + 7) / 8) {
        case 5: return launch_qb_mt<5>(dims, buf, flags, bwd, st);
        case 6: return launch_qb_mt<6>(dims, buf, flags, bwd, st);
        case 7: return launch_qb_mt<7>(dims, buf, flags, bwd, st);
        case 8: return launch_qb_mt<8>(dims, buf, flags, bwd, st);
        default: return false;
    }
}

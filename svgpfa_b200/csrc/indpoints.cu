// Batched per-(trial, latent) M x M work, M <= 64, one CTA per matrix, everything staged in
// shared memory:
//   kzz_chol_kernel       Kzz = kappa(Z,Z)+reg I, L = chol(Kzz), Li = L^-1, sum log L_ii
//   indpoints_fwd_kernel  Ls, X = Li Ls, c = Li m, alpha = Li^T c, KL_rk
//   indpoints_bwd_kernel  adjoints through alpha, c, X, KL and the Cholesky factorisation
// Reference arithmetic: stats/kernelsMatricesStore.py:107-138, utils/miscUtils.py:135-155,209-216,
// stats/klDivergence.py:31-44; adjoints per SURVEY.md Appendix A (the reference uses autograd).
#include "common.cuh"

namespace {

constexpr int IP_THREADS = 128;

__device__ __forceinline__ int ld_of(int M) { return M | 1; }   // odd leading dimension: conflict-free columns

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(64) kzz_chol_kernel(svgpfa_dims dm, svgpfa_buffers bf) {
    extern __shared__ double sm[];
    const int r = blockIdx.x, k = blockIdx.y;
    const svgpfa_latent_desc ds = bf.desc[k];
    const int M = ds.M, ld = ld_of(M);
    double* A = sm;                 // M x ld : Kzz -> L
    double* B = A + M * ld;         // M x ld : Li
    double* zs = B + M * ld;        // M
    double* dinv = zs + M;          // M
    const int tid = threadIdx.x, T = blockDim.x;
    const KConst kc = make_kconst(ds, bf.theta, bf.kscale, k);
    const double* z = bf.Z + (size_t)dm.R * ds.moff + (size_t)r * M;
    for (int i = tid; i < M; i += T) zs[i] = z[i];
    __syncthreads();
    for (int idx = tid; idx < M * M; idx += T) {
        const int i = idx / M, j = idx - i * M;
        if (j <= i) A[i * ld + j] = kappa_val(kc, zs[i] - zs[j]) + (i == j ? dm.reg : 0.0);
    }
    __syncthreads();
    // left-looking Cholesky, thread i owns row i
    bool bad = false;
    for (int j = 0; j < M; ++j) {
        double s = 0.0;
        if (tid >= j && tid < M) {
            s = A[tid * ld + j];
            for (int p = 0; p < j; ++p) s -= A[tid * ld + p] * A[j * ld + p];
        }
        if (tid == j) {
            if (!(s > 0.0)) bad = true;
            const double dg = sqrt(s);
            A[j * ld + j] = dg;
            dinv[j] = 1.0 / dg;
        }
        __syncthreads();
        if (tid > j && tid < M) A[tid * ld + j] = s * dinv[j];
        __syncthreads();
    }
    if (bad) {
        if (atomicCAS(bf.info, 0, SVGPFA_INFO_NOT_PD) == 0) { bf.info[1] = r; bf.info[2] = k; }
    }
    // Li by forward substitution, thread c owns column c
    if (tid < M) {
        const int c = tid;
        for (int i = 0; i < M; ++i) {
            double s = (i == c) ? 1.0 : 0.0;
            for (int p = c; p < i; ++p) s -= A[i * ld + p] * B[p * ld + c];
            B[i * ld + c] = (i < c) ? 0.0 : s * dinv[i];
        }
    }
    __syncthreads();
    double* Lg = bf.L + (size_t)r * dm.MM + ds.mmoff;
    double* Lig = bf.Li + (size_t)r * dm.MM + ds.mmoff;
    for (int idx = tid; idx < M * M; idx += T) {
        const int i = idx / M, j = idx - i * M;
        Lg[idx] = (j <= i) ? A[i * ld + j] : 0.0;
        Lig[idx] = (j <= i) ? B[i * ld + j] : 0.0;
    }
    if (tid == 0) {
        double s = 0.0;
        for (int i = 0; i < M; ++i) s += log(A[i * ld + i]);
        bf.logdetL[(size_t)r * dm.K + k] = s;
    }
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(IP_THREADS) indpoints_fwd_kernel(svgpfa_dims dm, svgpfa_buffers bf) {
    extern __shared__ double sm[];
    __shared__ double red[32];
    const int r = blockIdx.x, k = blockIdx.y;
    const svgpfa_latent_desc ds = bf.desc[k];
    const int M = ds.M, ld = ld_of(M);
    double* A = sm;                 // Li
    double* B = A + M * ld;         // Ls
    double* mv = B + M * ld;        // m
    double* cv = mv + M;            // c
    const int tid = threadIdx.x, T = blockDim.x;
    const double* Lig = bf.Li + (size_t)r * dm.MM + ds.mmoff;
    const double* cvec = bf.cholvec + (size_t)dm.R * ds.poff + (size_t)r * ds.P;
    const double* mg = bf.m + (size_t)dm.R * ds.moff + (size_t)r * M;
    for (int idx = tid; idx < M * M; idx += T) {
        const int i = idx / M, j = idx - i * M;
        A[i * ld + j] = Lig[idx];
        B[i * ld + j] = (j <= i) ? cvec[i * (i + 1) / 2 + j] : 0.0;
    }
    for (int i = tid; i < M; i += T) mv[i] = mg[i];
    __syncthreads();
    double* Xg = bf.X + (size_t)r * dm.MM + ds.mmoff;
    double part = 0.0;
    for (int idx = tid; idx < M * M; idx += T) {
        const int i = idx / M, j = idx - i * M;
        double s = 0.0;
        if (j <= i)
            for (int p = j; p <= i; ++p) s += A[i * ld + p] * B[p * ld + j];
        Xg[idx] = s;
        part += s * s;
    }
    for (int i = tid; i < M; i += T) {
        double s = 0.0;
        for (int p = 0; p <= i; ++p) s += A[i * ld + p] * mv[p];
        cv[i] = s;
        part += s * s - 2.0 * log(fabs(B[i * ld + i]));
    }
    __syncthreads();
    const size_t vo = (size_t)r * dm.KM + ds.moff;
    for (int j = tid; j < M; j += T) {
        double s = 0.0;
        for (int i = j; i < M; ++i) s += A[i * ld + j] * cv[i];
        bf.alpha[vo + j] = s;
        bf.c[vo + j] = cv[j];
    }
    const double tot = block_sum(part, red);
    if (tid == 0)
        bf.kl_rk[(size_t)r * dm.K + k] = 0.5 * (tot + 2.0 * bf.logdetL[(size_t)r * dm.K + k] - (double)M);
}

// ------------------------------------------------------------------------------------------
// out(i,j) for all (i,j) with pred(i,j): one output element per thread-iteration.
template <class F>
__device__ __forceinline__ void for_each_ij(int M, bool lower_only, F f) {
    for (int idx = threadIdx.x; idx < M * M; idx += blockDim.x) {
        const int i = idx / M, j = idx - i * M;
        if (!lower_only || j <= i) f(i, j);
    }
}

__global__ void __launch_bounds__(IP_THREADS) indpoints_bwd_kernel(svgpfa_dims dm, svgpfa_buffers bf, uint32_t flags) {
    extern __shared__ double sm[];
    __shared__ double red[32];
    const int r = blockIdx.x, k = blockIdx.y;
    const svgpfa_latent_desc ds = bf.desc[k];
    const int M = ds.M, ld = ld_of(M), MS = M * ld;
    const bool need_post = flags & SVGPFA_GRAD_POSTERIOR;
    const bool need_kz = flags & (SVGPFA_GRAD_KERNEL | SVGPFA_GRAD_INDLOCS);
    double* Lm = sm;             // L
    double* Li = Lm + MS;        // L^-1
    double* X = Li + MS;         // L^-1 Ls
    double* B3 = X + MS;
    double* B4 = B3 + MS;
    double* B5 = B4 + MS;
    double* al = B5 + MS;        // alpha
    double* cv = al + M;         // c
    double* yv = cv + M;         // Li abar
    double* mb = yv + M;         // mbar
    double* ab = mb + M;         // abar total
    double* zs = ab + M;         // z
    const int tid = threadIdx.x, T = blockDim.x;
    const size_t mo = (size_t)r * dm.MM + ds.mmoff, vo = (size_t)r * dm.KM + ds.moff;
    for (int idx = tid; idx < M * M; idx += T) {
        const int i = idx / M, j = idx - i * M;
        Lm[i * ld + j] = bf.L[mo + idx];
        Li[i * ld + j] = bf.Li[mo + idx];
        X[i * ld + j] = bf.X[mo + idx];
        // A_q is stored lower; mirror it
        B3[i * ld + j] = (j <= i) ? bf.A_q[mo + idx] : bf.A_q[mo + (size_t)j * M + i];
    }
    const double* zg = bf.Z + (size_t)dm.R * ds.moff + (size_t)r * M;
    for (int i = tid; i < M; i += T) {
        al[i] = bf.alpha[vo + i];
        cv[i] = bf.c[vo + i];
        ab[i] = bf.abar_q[vo + i] + bf.abar_spk[vo + i];
        zs[i] = zg[i];
    }
    __syncthreads();
    // y = Li abar ; cbar = y - c
    for (int i = tid; i < M; i += T) {
        double s = 0.0;
        for (int p = 0; p <= i; ++p) s += Li[i * ld + p] * ab[p];
        yv[i] = s;
    }
    // Xbar = tril(2 A X) - X   -> B4
    for_each_ij(M, false, [&](int i, int j) {
        double s = 0.0;
        if (j <= i) {
            for (int p = j; p < M; ++p) s += B3[i * ld + p] * X[p * ld + j];
            s = 2.0 * s - X[i * ld + j];
        }
        B4[i * ld + j] = s;
    });
    __syncthreads();
    // mbar = Li^T cbar
    for (int j = tid; j < M; j += T) {
        double s = 0.0;
        for (int i = j; i < M; ++i) s += Li[i * ld + j] * (yv[i] - cv[i]);
        mb[j] = s;
    }
    // T = Li^T tril(Xbar) -> B5  (full matrix needed for T X^T)
    for_each_ij(M, false, [&](int i, int j) {
        double s = 0.0;
        for (int p = (i > j ? i : j); p < M; ++p) s += Li[p * ld + i] * B4[p * ld + j];
        B5[i * ld + j] = s;
    });
    __syncthreads();
    if (need_post) {
        double* gm = bf.gm + (size_t)dm.R * ds.moff + (size_t)r * M;
        for (int i = tid; i < M; i += T) gm[i] = mb[i];
        double* gcv = bf.gcholvec + (size_t)dm.R * ds.poff + (size_t)r * ds.P;
        const double* cvec = bf.cholvec + (size_t)dm.R * ds.poff + (size_t)r * ds.P;
        for_each_ij(M, true, [&](int i, int j) {
            const int p = i * (i + 1) / 2 + j;
            gcv[p] = B5[i * ld + j] + (i == j ? 1.0 / cvec[p] : 0.0);
        });
    }
    if (!need_kz) return;
    // B4 = X^T A
    for_each_ij(M, false, [&](int i, int j) {
        double s = 0.0;
        for (int p = i; p < M; ++p) s += X[p * ld + i] * B3[p * ld + j];
        B4[i * ld + j] = s;
    });
    __syncthreads();
    // B3 = E2 = X (X^T A) - A   (in place: each thread touches only its own element of B3)
    for_each_ij(M, false, [&](int i, int j) {
        double s = 0.0;
        for (int p = 0; p <= i; ++p) s += X[i * ld + p] * B4[p * ld + j];
        B3[i * ld + j] = s - B3[i * ld + j];
    });
    __syncthreads();
    // Lbar (lower) -> B4 = -2 Li^T E2 - alpha y^T - diag(1/L_ii) - T X^T - mbar c^T
    for_each_ij(M, false, [&](int i, int j) {
        double s = 0.0;
        if (j <= i) {
            double q = 0.0;
            for (int p = i; p < M; ++p) q += Li[p * ld + i] * B3[p * ld + j];
            double t = 0.0;
            for (int p = 0; p <= j; ++p) t += B5[i * ld + p] * X[j * ld + p];
            s = -2.0 * q - al[i] * yv[j] - t - mb[i] * cv[j];
            if (i == j) s -= 1.0 / Lm[i * ld + i];
        }
        B4[i * ld + j] = s;
    });
    __syncthreads();
    // P = Phi(L^T Lbar) -> B3 lower, then symmetrise: S = P + P^T
    for_each_ij(M, true, [&](int i, int j) {
        double s = 0.0;
        for (int p = i; p < M; ++p) s += Lm[p * ld + i] * B4[p * ld + j];
        B3[i * ld + j] = s;            // diagonal: P_ii = s/2, S_ii = s
    });
    __syncthreads();
    for_each_ij(M, false, [&](int i, int j) {
        if (j > i) B3[i * ld + j] = B3[j * ld + i];
    });
    __syncthreads();
    // U1 = S Li -> B4
    for_each_ij(M, false, [&](int i, int j) {
        double s = 0.0;
        for (int p = j; p < M; ++p) s += B3[i * ld + p] * Li[p * ld + j];
        B4[i * ld + j] = s;
    });
    __syncthreads();
    // Kbar = 0.5 Li^T U1 -> B5
    for_each_ij(M, false, [&](int i, int j) {
        double s = 0.0;
        for (int p = i; p < M; ++p) s += Li[p * ld + i] * B4[p * ld + j];
        B5[i * ld + j] = 0.5 * s;
    });
    __syncthreads();
    // dZ_i = 2 sum_j Kbar_ij dkappa/ddelta(z_i - z_j);  dtheta = sum_ij Kbar_ij dkappa/dtheta
    const KConst kc = make_kconst(ds, bf.theta, bf.kscale, k);
    double t0 = 0.0, t1 = 0.0;
    double* gZ = bf.gZ + (size_t)dm.R * ds.moff + (size_t)r * M;
    for (int i = tid; i < M; i += T) {
        double dz = 0.0;
        for (int j = 0; j < M; ++j) {
            double kv, dkd, d0, d1;
            kappa_grad(kc, zs[i] - zs[j], kv, dkd, d0, d1);
            const double kb = B5[i * ld + j];
            dz += kb * dkd;
            t0 += kb * d0;
            t1 += kb * d1;
        }
        if (flags & SVGPFA_GRAD_INDLOCS) gZ[i] = 2.0 * dz + bf.dz_acc[vo + i];
    }
    if (flags & SVGPFA_GRAD_KERNEL) {
        const double s0 = block_sum(t0, red);
        const double s1 = block_sum(t1, red);
        if (tid == 0) {
            double* dth = bf.dth_part + (size_t)r * dm.TH + ds.thoff;
            dth[0] += s0;
            if (ds.nth > 1) dth[1] += s1;
        }
    }
}

size_t ip_smem(int Mmax, int nmat, int nvec) {
    const int ld = Mmax | 1;
    return sizeof(double) * ((size_t)nmat * Mmax * ld + (size_t)nvec * Mmax);
}

}  // namespace

extern "C" int svgpfa_kzz_chol_fwd(const svgpfa_dims* dims, const svgpfa_buffers* buf, void* stream) {
    if (!dims || !buf || dims->Mmax > SVGPFA_MAX_M || dims->Mmax < 1) return svgpfa_set_error(SVGPFA_E_ARG, "kzz_chol_fwd", cudaSuccess);
    if (dims->R == 0) return SVGPFA_OK;
    const size_t smem = ip_smem(dims->Mmax, 2, 2);
    cudaFuncSetAttribute(kzz_chol_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kzz_chol_kernel<<<dim3(dims->R, dims->K), 64, smem, (cudaStream_t)stream>>>(*dims, *buf);
    SVGPFA_CHECK_LAUNCH("kzz_chol_fwd");
    return SVGPFA_OK;
}

extern "C" int svgpfa_indpoints_fwd(const svgpfa_dims* dims, const svgpfa_buffers* buf, void* stream) {
    if (!dims || !buf || dims->Mmax > SVGPFA_MAX_M) return svgpfa_set_error(SVGPFA_E_ARG, "indpoints_fwd", cudaSuccess);
    if (dims->R == 0) return SVGPFA_OK;
    const size_t smem = ip_smem(dims->Mmax, 2, 2);
    cudaFuncSetAttribute(indpoints_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    indpoints_fwd_kernel<<<dim3(dims->R, dims->K), IP_THREADS, smem, (cudaStream_t)stream>>>(*dims, *buf);
    SVGPFA_CHECK_LAUNCH("indpoints_fwd");
    return SVGPFA_OK;
}

extern "C" int svgpfa_indpoints_bwd(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, void* stream) {
    if (!dims || !buf || dims->Mmax > SVGPFA_MAX_M) return svgpfa_set_error(SVGPFA_E_ARG, "indpoints_bwd", cudaSuccess);
    if (dims->R == 0) return SVGPFA_OK;
    const size_t smem = ip_smem(dims->Mmax, 6, 6);
    cudaFuncSetAttribute(indpoints_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(indpoints_bwd_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    indpoints_bwd_kernel<<<dim3(dims->R, dims->K), IP_THREADS, smem, (cudaStream_t)stream>>>(*dims, *buf, flags);
    SVGPFA_CHECK_LAUNCH("indpoints_bwd");
    return SVGPFA_OK;
}

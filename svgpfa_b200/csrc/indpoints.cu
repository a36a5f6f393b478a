// Batched per-(trial, latent) M x M work, M <= 64:
//   kzz_chol_warp_kernel (M <= 32, warp per matrix), kzz_chol_pair_kernel (M <= 64, two warps per matrix)
//                         Kzz = kappa(Z,Z)+reg I, L = chol(Kzz), Li = L^-1, sum log L_ii; fused with the stage below
//                         inside svgpfa_elbo_grad
//   indpoints_fwd_warp_kernel (M <= 32), indpoints_fwd_kernel      Ls, X = Li Ls, c = Li m, alpha = Li^T c, KL_rk
//   indpoints_bwd_mma_kernel  adjoints through alpha, c, X, KL and the Cholesky factorisation (M <= 64)
// Reference arithmetic: stats/kernelsMatricesStore.py:107-138, utils/miscUtils.py:135-155,209-216,
// stats/klDivergence.py:31-44; adjoints per SURVEY.md Appendix A (the reference uses autograd).
#include "common.cuh"

namespace {

constexpr int IP_THREADS = 128;

__device__ __forceinline__ int ld_of(int M) { return M | 1; }   // odd leading dimension: conflict-free columns

// ------------------------------------------------------------------------------------------
// M <= 32: one WARP per (trial, latent) matrix, rows in registers.
// The first version (one thread per row or column, the matrix in shared memory: two loads per FMA) needed ~110 k cycles
// per matrix (ncu: 45 % fixed-latency "wait" stalls, 36 % of the samples in the column-wise inverse).
// Here lane i owns ROW i of the trailing matrix in registers: a right-looking Cholesky whose column broadcasts are warp
// shuffles, written with a rotating register window (after column j is eliminated a[k-1] <- a[k] - l_ij l_kj, so the
// loop over j stays rolled and the code small), followed by the column-wise inverse with lane c owning column c of
// L^-1 in registers and the entries of L read as warp-uniform (broadcast) shared-memory loads.  No block barriers.
constexpr int KC_WARPS = 4;
constexpr int KC_LD = 33;                         // odd: lane <-> row accesses are bank-conflict free
constexpr int KC_COL = 64;                        // column broadcast buffer: entries 32..63 stay zero (the rows past the edge)
constexpr int KC_WSM = 32 * KC_LD + 32 + KC_COL;  // doubles of shared memory per warp: matrix + 1 / L_jj + column buffer

// 8 columns of the factorisation on a window of LEN columns (LEN = 32 - j0).  Column j of L reaches the other lanes
// through a 64-entry shared-memory buffer read with warp-uniform (broadcast) LDS.64: one instruction per entry where a
// 64-bit shuffle is two, and the shuffles were the kernel's busiest pipe (608 per matrix).
template <int LEN>
__device__ __forceinline__ void chol_stage(double (&a)[32], int j0, int lane, int M, double* __restrict__ A,
                                           double* __restrict__ dinv, double* __restrict__ colb, bool& bad, double& mydiag) {
#pragma unroll 1
    for (int j = j0; j < j0 + 8; ++j) {
        const double d = __shfl_sync(0xffffffffu, a[0], j);
        if (j < M && !(d > 0.0)) bad = true;
        const double dg = sqrt(d), inv = 1.0 / dg;
        const double lij = lane == j ? dg : (lane > j ? a[0] * inv : 0.0);
        if (lane == j) mydiag = dg;
        A[lane * KC_LD + j] = lij;
        if (lane == 0) dinv[j] = inv;
        colb[lane] = lij;
        __syncwarp();
        const double* cj = colb + j;
#pragma unroll
        for (int kk = 1; kk < LEN; ++kk)
            a[kk - 1] = fma(-lij, cj[kk], a[kk]);        // columns past the matrix edge carry garbage, never read
        __syncwarp();                                    // the buffer is rewritten by the next column
    }
}

template <bool LOAD_LI>
__device__ __forceinline__ void ipfwd_warp_body(const svgpfa_dims& dm, const svgpfa_buffers& bf, int r, int k,
                                                const svgpfa_latent_desc& ds, double* __restrict__ A, double* __restrict__ vec,
                                                int lane);

template <bool FUSE_FWD>
__global__ void __launch_bounds__(32 * KC_WARPS) kzz_chol_warp_kernel(svgpfa_dims dm, svgpfa_buffers bf, int nprob) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int prob = blockIdx.x * KC_WARPS + warp;
    if (prob >= nprob) return;                        // whole warps leave; there is no block barrier below
    const int rl = prob / dm.K, k = prob - rl * dm.K, r = dm.r0 + rl;
    const svgpfa_latent_desc ds = bf.desc[k];
    const int M = ds.M;
    double* A = sm + (size_t)warp * KC_WSM;
    double* dinv = A + 32 * KC_LD;
    double* colb = dinv + 32;
    colb[32 + lane] = 0.0;
    const KConst kc = make_kconst(ds, bf.theta, bf.kscale, k);
    const double zi = lane < M ? bf.Z[(size_t)dm.R * ds.moff + (size_t)r * M + lane] : 0.0;
    // Kzz (lower triangle) -> A; rows / columns >= M are those of the identity.  Rows t and 30 - t together hold 32
    // lower-triangle entries, so 17 passes cover the 528 entries with (almost) every lane busy.
    for (int t = 0; t < 17; ++t) {
        int i, j;
        if (t < 15) { i = lane <= t ? t : 30 - t; j = lane <= t ? lane : lane - t - 1; }
        else if (t == 15) { i = 15; j = lane; }
        else { i = 31; j = lane; }
        const double z1 = __shfl_sync(0xffffffffu, zi, i), z2 = __shfl_sync(0xffffffffu, zi, j & 31);
        if (j <= i) {
            double v = i == j ? 1.0 : 0.0;
            if (i < M && j < M) v = kappa_val(kc, z1 - z2) + (i == j ? dm.reg : 0.0);
            A[i * KC_LD + j] = v;
        }
    }
    __syncwarp();
    double a[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) a[j] = j <= lane ? A[lane * KC_LD + j] : 0.0;
    __syncwarp();
    bool bad = false;
    double mydiag = 1.0;
    chol_stage<32>(a, 0, lane, M, A, dinv, colb, bad, mydiag);
    chol_stage<24>(a, 8, lane, M, A, dinv, colb, bad, mydiag);
    chol_stage<16>(a, 16, lane, M, A, dinv, colb, bad, mydiag);
    chol_stage<8>(a, 24, lane, M, A, dinv, colb, bad, mydiag);
    __syncwarp();
    if (bad && lane == 0) {
        if (atomicCAS(bf.info, 0, SVGPFA_INFO_NOT_PD) == 0) { bf.info[1] = r; bf.info[2] = k; }
    }
    double* Lg = bf.L + (size_t)r * dm.MM + ds.mmoff;
    double* Lig = bf.Li + (size_t)r * dm.MM + ds.mmoff;
    for (int idx = lane; idx < M * M; idx += 32) {
        const int i = idx / M, j = idx - i * M;
        Lg[idx] = A[i * KC_LD + j];                   // zeros above the diagonal
    }
    const double ld = warp_sum(lane < M ? log(mydiag) : 0.0);
    if (lane == 0) bf.logdetL[(size_t)r * dm.K + k] = ld;
    // L^-1: lane c owns column c;  x_i = (delta_ic - sum_{p < i} L_ip x_p) / L_ii  (x_p = 0 for p < c by itself)
    double x[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        double s0 = i == lane ? 1.0 : 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int p = 0; p < i; ++p) {
            const double l = A[i * KC_LD + p];
            if ((p & 3) == 0) s0 = fma(-l, x[p], s0);
            else if ((p & 3) == 1) s1 = fma(-l, x[p], s1);
            else if ((p & 3) == 2) s2 = fma(-l, x[p], s2);
            else s3 = fma(-l, x[p], s3);
        }
        x[i] = ((s0 + s1) + (s2 + s3)) * dinv[i];
    }
    __syncwarp();                                     // every lane is done reading L
#pragma unroll
    for (int i = 0; i < 32; ++i) A[i * KC_LD + lane] = x[i];
    __syncwarp();
    for (int idx = lane; idx < M * M; idx += 32) {
        const int i = idx / M, j = idx - i * M;
        Lig[idx] = j <= i ? A[i * KC_LD + j] : 0.0;
    }
    if (FUSE_FWD) {
        // X = Li Ls, c, alpha, KL straight from the L^-1 that is still in this warp's tile: saves the launch and the
        // re-read of Li (3.3 GB at config #5); logdetL is read back by the same lane that wrote it
        __syncwarp();
        ipfwd_warp_body<false>(dm, bf, r, k, ds, A, dinv, lane);
    }
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(IP_THREADS) indpoints_fwd_kernel(svgpfa_dims dm, svgpfa_buffers bf) {
    extern __shared__ double sm[];
    __shared__ double red[32];
    const int r = dm.r0 + blockIdx.x, k = blockIdx.y;
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
// M <= 32: the same work with one WARP per (trial, latent), like kzz_chol_warp_kernel: Li staged in shared memory
// (coalesced), lane j owns column j of Ls in registers, X[i][j] = sum_p Li[i][p] Ls[p][j] with the entries of Li read
// as warp-uniform (broadcast) loads and row i of X written straight from the registers (one coalesced store per row);
// c = Li m with lane = row, alpha = Li^T c with lane = column.  No block barriers.
// A: this warp's [32][KC_LD] tile; LOAD_LI: fetch Li from global memory (otherwise the tile already holds it -- the fused
// Cholesky kernel below leaves L^-1 there, with unit diagonal in the padding rows, which the zero-padded Ls / m ignore)
template <bool LOAD_LI>
__device__ __forceinline__ void ipfwd_warp_body(const svgpfa_dims& dm, const svgpfa_buffers& bf, int r, int k,
                                                const svgpfa_latent_desc& ds, double* __restrict__ A, double* __restrict__ vec,
                                                int lane) {
    const int M = ds.M;
    const size_t mo = (size_t)r * dm.MM + ds.mmoff, vo = (size_t)r * dm.KM + ds.moff;
    if (LOAD_LI) {
        const double* Lig = bf.Li + mo;
        for (int i = 0; i < 32; ++i) A[i * KC_LD + lane] = (i < M && lane < M) ? Lig[(size_t)i * M + lane] : 0.0;
    }
    vec[lane] = lane < M ? bf.m[(size_t)dm.R * ds.moff + (size_t)r * M + lane] : 0.0;
    const double* cvec = bf.cholvec + (size_t)dm.R * ds.poff + (size_t)r * ds.P;
    double ls[32];                                    // column `lane` of Ls (row-major tril vector, miscUtils.py:135-139)
#pragma unroll
    for (int p = 0; p < 32; ++p) ls[p] = (p < M && lane <= p) ? cvec[p * (p + 1) / 2 + lane] : 0.0;
    const double ldiag = lane < M ? cvec[lane * (lane + 1) / 2 + lane] : 1.0;
    __syncwarp();
    double part = 0.0;
    double* Xg = bf.X + mo;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int p = 0; p <= i; ++p) {
            const double l = A[i * KC_LD + p];
            if ((p & 3) == 0) s0 = fma(l, ls[p], s0);
            else if ((p & 3) == 1) s1 = fma(l, ls[p], s1);
            else if ((p & 3) == 2) s2 = fma(l, ls[p], s2);
            else s3 = fma(l, ls[p], s3);
        }
        const double x = (s0 + s1) + (s2 + s3);       // 0 above the diagonal by itself (ls[p] = 0 for p < lane)
        if (i < M && lane < M) Xg[(size_t)i * M + lane] = x;
        part = fma(x, x, part);
    }
    double c = 0.0;                                   // c = Li m, lane = row (Li is stored with zeros above the diagonal)
#pragma unroll 8
    for (int p = 0; p < 32; ++p) c = fma(A[lane * KC_LD + p], vec[p], c);
    part += c * c - 2.0 * log(fabs(ldiag));
    __syncwarp();
    vec[lane] = c;
    __syncwarp();
    double al = 0.0;                                  // alpha = Li^T c, lane = column
#pragma unroll 8
    for (int i = 0; i < 32; ++i) al = fma(A[i * KC_LD + lane], vec[i], al);
    if (lane < M) {
        bf.alpha[vo + lane] = al;
        bf.c[vo + lane] = c;
    }
    const double tot = warp_sum(part);
    if (lane == 0) bf.kl_rk[(size_t)r * dm.K + k] = 0.5 * (tot + 2.0 * bf.logdetL[(size_t)r * dm.K + k] - (double)M);
}

__global__ void __launch_bounds__(32 * KC_WARPS) indpoints_fwd_warp_kernel(svgpfa_dims dm, svgpfa_buffers bf, int nprob) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int prob = blockIdx.x * KC_WARPS + warp;
    if (prob >= nprob) return;                        // whole warps leave; there is no block barrier below
    const int rl = prob / dm.K, k = prob - rl * dm.K, r = dm.r0 + rl;
    const svgpfa_latent_desc ds = bf.desc[k];
    double* A = sm + (size_t)warp * KC_WSM;           // Li, rows / columns >= M zero
    ipfwd_warp_body<true>(dm, bf, r, k, ds, A, A + 32 * KC_LD, lane);
}

// ------------------------------------------------------------------------------------------
// 32 < M <= 64: the same scheme with TWO warps per matrix (CTA = 64 threads = one matrix): thread t owns row t of the
// trailing matrix in a rotating window of 64 registers, then column t of L^-1, then column t of Ls.  What the warp
// version passes through shuffles goes through shared memory here: every thread posts its entry of the current
// column UNNORMALISED, one barrier, and everybody reads the pivot d and the column from there,
//     a[k-1] <- a[k] - (a_i / d) a_k        (= a[k] - l_ij l_kj),
// so a column costs one barrier (the column buffer is double-buffered).  Replaces a one-thread-per-row kernel that
// read both factors of every product from shared memory (M = 64, R = 1000: 4.7 + 1.9 ms for the Cholesky and the
// inducing-point forward stage, 2.7 ms now).
constexpr int KP_N = 64, KP_LD = 65, KP_COL = 128;

template <int LEN>
__device__ __forceinline__ void pair_stage(double (&a)[KP_N], int j0, int t, int M, double* __restrict__ A,
                                           double* __restrict__ dinv, double* __restrict__ colb, bool& bad, double& mydiag) {
#pragma unroll 1
    for (int j = j0; j < j0 + 8; ++j) {
        double* cb = colb + (j & 1) * KP_COL;
        cb[t] = t >= j ? a[0] : 0.0;
        __syncthreads();
        const double d = cb[j];
        if (j < M && !(d > 0.0)) bad = true;
        const double dg = sqrt(d), inv = 1.0 / dg;
        A[t * KP_LD + j] = t == j ? dg : (t > j ? a[0] * inv : 0.0);
        if (t == j) mydiag = dg;
        if (t == 0) dinv[j] = inv;
        const double f = t > j ? a[0] * (inv * inv) : 0.0;
        const double* cj = cb + j;
#pragma unroll
        for (int kk = 1; kk < LEN; ++kk) a[kk - 1] = fma(-f, cj[kk], a[kk]);     // columns past the edge: cj[] = 0
    }
}

template <bool FUSE_FWD>
__global__ void __launch_bounds__(KP_N) kzz_chol_pair_kernel(svgpfa_dims dm, svgpfa_buffers bf) {
    __shared__ double A[KP_N * KP_LD];
    __shared__ double dinv[KP_N], colb[2 * KP_COL], mvec[KP_N], cvs[KP_N], red[32];
    const int t = threadIdx.x;
    const int r = dm.r0 + blockIdx.x, k = blockIdx.y;
    const svgpfa_latent_desc ds = bf.desc[k];
    const int M = ds.M;
    const KConst kc = make_kconst(ds, bf.theta, bf.kscale, k);
    for (int i = t; i < 2 * KP_COL; i += KP_N) colb[i] = 0.0;
    mvec[t] = t < M ? bf.Z[(size_t)dm.R * ds.moff + (size_t)r * M + t] : 0.0;       // z for now
    __syncthreads();
    // Kzz (lower triangle), rows / columns >= M those of the identity; the 2080 entries dealt evenly to the threads
    for (int idx = t; idx < KP_N * (KP_N + 1) / 2; idx += KP_N) {
        int i = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
        while (i * (i + 1) / 2 > idx) --i;
        while ((i + 1) * (i + 2) / 2 <= idx) ++i;
        const int j = idx - i * (i + 1) / 2;
        double v = i == j ? 1.0 : 0.0;
        if (i < M && j < M) v = kappa_val(kc, mvec[i] - mvec[j]) + (i == j ? dm.reg : 0.0);
        A[i * KP_LD + j] = v;
    }
    __syncthreads();
    double a[KP_N];
#pragma unroll
    for (int j = 0; j < KP_N; ++j) a[j] = j <= t ? A[t * KP_LD + j] : 0.0;
    __syncthreads();
    bool bad = false;
    double mydiag = 1.0;
    pair_stage<64>(a, 0, t, M, A, dinv, colb, bad, mydiag);
    pair_stage<56>(a, 8, t, M, A, dinv, colb, bad, mydiag);
    pair_stage<48>(a, 16, t, M, A, dinv, colb, bad, mydiag);
    pair_stage<40>(a, 24, t, M, A, dinv, colb, bad, mydiag);
    pair_stage<32>(a, 32, t, M, A, dinv, colb, bad, mydiag);
    pair_stage<24>(a, 40, t, M, A, dinv, colb, bad, mydiag);
    pair_stage<16>(a, 48, t, M, A, dinv, colb, bad, mydiag);
    pair_stage<8>(a, 56, t, M, A, dinv, colb, bad, mydiag);
    __syncthreads();
    if (bad && t == 0) {
        if (atomicCAS(bf.info, 0, SVGPFA_INFO_NOT_PD) == 0) { bf.info[1] = r; bf.info[2] = k; }
    }
    const size_t mo = (size_t)r * dm.MM + ds.mmoff, vo = (size_t)r * dm.KM + ds.moff;
    double* Lg = bf.L + mo;
    double* Lig = bf.Li + mo;
    for (int idx = t; idx < M * M; idx += KP_N) {
        const int i = idx / M, j = idx - i * M;
        Lg[idx] = A[i * KP_LD + j];                   // zeros above the diagonal
    }
    const double ld = block_sum(t < M ? log(mydiag) : 0.0, red);
    if (t == 0) bf.logdetL[(size_t)r * dm.K + k] = ld;
    // L^-1: thread c owns column c;  x_i = (delta_ic - sum_{p < i} L_ip x_p) / L_ii
    double x[KP_N];
#pragma unroll
    for (int i = 0; i < KP_N; ++i) {
        double s0 = i == t ? 1.0 : 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int p = 0; p < i; ++p) {
            const double l = A[i * KP_LD + p];
            if ((p & 3) == 0) s0 = fma(-l, x[p], s0);
            else if ((p & 3) == 1) s1 = fma(-l, x[p], s1);
            else if ((p & 3) == 2) s2 = fma(-l, x[p], s2);
            else s3 = fma(-l, x[p], s3);
        }
        x[i] = ((s0 + s1) + (s2 + s3)) * dinv[i];
    }
    __syncthreads();                                  // every thread is done reading L
#pragma unroll
    for (int i = 0; i < KP_N; ++i) A[i * KP_LD + t] = x[i];
    __syncthreads();
    for (int idx = t; idx < M * M; idx += KP_N) {
        const int i = idx / M, j = idx - i * M;
        Lig[idx] = j <= i ? A[i * KP_LD + j] : 0.0;
    }
    if (!FUSE_FWD) return;
    // X = Li Ls, c = Li m, alpha = Li^T c, KL (indpoints_fwd): thread t owns column t of Ls in registers
    mvec[t] = t < M ? bf.m[(size_t)dm.R * ds.moff + (size_t)r * M + t] : 0.0;
    const double* cvec = bf.cholvec + (size_t)dm.R * ds.poff + (size_t)r * ds.P;
    double ls[KP_N];
#pragma unroll
    for (int p = 0; p < KP_N; ++p) ls[p] = (p < M && t <= p) ? cvec[p * (p + 1) / 2 + t] : 0.0;
    const double ldiag = t < M ? cvec[t * (t + 1) / 2 + t] : 1.0;
    __syncthreads();
    double part = 0.0;
    double* Xg = bf.X + mo;
#pragma unroll
    for (int i = 0; i < KP_N; ++i) {
        double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
        for (int p = 0; p <= i; ++p) {
            const double l = A[i * KP_LD + p];
            if ((p & 3) == 0) s0 = fma(l, ls[p], s0);
            else if ((p & 3) == 1) s1 = fma(l, ls[p], s1);
            else if ((p & 3) == 2) s2 = fma(l, ls[p], s2);
            else s3 = fma(l, ls[p], s3);
        }
        const double xv = (s0 + s1) + (s2 + s3);      // 0 above the diagonal by itself (ls[p] = 0 for p < t)
        if (i < M && t < M) Xg[(size_t)i * M + t] = xv;
        part = fma(xv, xv, part);
    }
    double c = 0.0;                                   // c = Li m, thread = row (padding rows: unit diagonal times m = 0)
#pragma unroll 8
    for (int p = 0; p < KP_N; ++p) c = fma(p <= t ? A[t * KP_LD + p] : 0.0, mvec[p], c);
    part += c * c - 2.0 * log(fabs(ldiag));
    cvs[t] = c;
    __syncthreads();
    double al = 0.0;                                  // alpha = Li^T c, thread = column
#pragma unroll 8
    for (int i = 0; i < KP_N; ++i) al = fma(i >= t ? A[i * KP_LD + t] : 0.0, cvs[i], al);
    if (t < M) {
        bf.alpha[vo + t] = al;
        bf.c[vo + t] = c;
    }
    const double tot = block_sum(part, red);
    if (t == 0) bf.kl_rk[(size_t)r * dm.K + k] = 0.5 * (tot + 2.0 * ld - (double)M);
}

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void dmma8(double& c0, double& c1, double a, double b) {
    asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
        : "+d"(c0), "+d"(c1)
        : "d"(a), "d"(b));
}

// ------------------------------------------------------------------------------------------
// Adjoints through alpha, c, X, the KL term and the Cholesky factorisation, nine M x M products on
// the FP64 tensor path with compile-time shapes.  What ncu showed of the round-1 kernel (one warp per 8x8 output tile,
// generic lambdas for the operands and the k ranges: 27 900 warp instructions per matrix for ~650 mma, issue slots 44 %
// busy, FP64 pipe 5 %): ~40 instructions per mma.  Here warp `it` owns the 8-row tile `it` of every product and all
// its column tiles (one A fragment per k-step shared by the MT column tiles), the k range of each (row, column) tile is
// a compile-time-unrolled predicated loop, and a CTA is MT warps: 10 500 instructions per matrix, 4 CTAs per SM at
// M = 32.  Matrices are stored full with explicit zeros outside their triangle and zero padding up to MP = 8 MT.
// ------------------------------------------------------------------------------------------
// acc(jt) = sum_{ks in [lo(jt), hi(jt)]} A(8 it + g, 4 ks + tg) B(4 ks + tg, 8 jt + g);  TA / TB: operand stored transposed
template <int MT, bool TA, bool TB, class FR, class FO>
__device__ __forceinline__ void ipb_product(const double* __restrict__ A, const double* __restrict__ B, int it, int lane,
                                            FR range, FO out) {
    constexpr int LD = 8 * MT + 4, KS = 2 * MT;
    const int g = lane >> 2, tg = lane & 3;
    double acc[MT][2];
    int lo[MT], hi[MT];
#pragma unroll
    for (int jt = 0; jt < MT; ++jt) {
        acc[jt][0] = acc[jt][1] = 0.0;
        range(it, jt, lo[jt], hi[jt]);                    // k-steps; hi < lo: the tile is not needed
    }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        const double a = TA ? A[(4 * ks + tg) * LD + 8 * it + g] : A[(8 * it + g) * LD + 4 * ks + tg];
#pragma unroll
        for (int jt = 0; jt < MT; ++jt) {
            if (ks >= lo[jt] && ks <= hi[jt]) {
                const double b = TB ? B[(8 * jt + g) * LD + 4 * ks + tg] : B[(4 * ks + tg) * LD + 8 * jt + g];
                dmma8(acc[jt][0], acc[jt][1], a, b);
            }
        }
    }
#pragma unroll
    for (int jt = 0; jt < MT; ++jt) {
        if (hi[jt] >= lo[jt]) {
            out(8 * it + g, 8 * jt + 2 * tg, acc[jt][0]);
            out(8 * it + g, 8 * jt + 2 * tg + 1, acc[jt][1]);
        }
    }
}

template <int MT>
__global__ void __launch_bounds__(32 * MT) indpoints_bwd_mma_kernel(svgpfa_dims dm, svgpfa_buffers bf, uint32_t flags) {
    constexpr int MP = 8 * MT, LD = MP + 4, MS = MP * LD, KS = 2 * MT, T = 32 * MT;
    extern __shared__ __align__(16) double sm[];
    __shared__ double red[32];
    const int r = dm.r0 + blockIdx.x, k = blockIdx.y;
    const svgpfa_latent_desc ds = bf.desc[k];
    const int M = ds.M;
    const bool need_post = flags & SVGPFA_GRAD_POSTERIOR;
    const bool need_kz = flags & (SVGPFA_GRAD_KERNEL | SVGPFA_GRAD_INDLOCS);
    double* Lm = sm;             // L
    double* Li = Lm + MS;        // L^-1
    double* X = Li + MS;         // L^-1 Ls
    double* B3 = X + MS;
    double* B4 = B3 + MS;
    double* B5 = B4 + MS;
    double* al = B5 + MS;        // alpha
    double* cv = al + MP;        // c
    double* yv = cv + MP;        // Li abar
    double* mb = yv + MP;        // mbar
    double* ab = mb + MP;        // abar total
    double* zs = ab + MP;        // z
    const int tid = threadIdx.x, lane = tid & 31, it = tid >> 5;
    const size_t mo = (size_t)r * dm.MM + ds.mmoff, vo = (size_t)r * dm.KM + ds.moff;
    const KConst kc = make_kconst(ds, bf.theta, bf.kscale, k);
    if (M == MP && (mo & 1) == 0) {                      // whole 16-byte chunks: asynchronous copies
        const unsigned s0 = (unsigned)__cvta_generic_to_shared(Lm);
        for (int c = tid; c < MP * (MP / 2); c += T) {
            const int i = c / (MP / 2), jj = c - i * (MP / 2);
            const size_t go = mo + (size_t)i * M + 2 * jj;
            const unsigned so = (unsigned)(i * LD + 2 * jj) * 8u;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0 + so), "l"(bf.L + go) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0 + MS * 8u + so), "l"(bf.Li + go) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0 + 2u * MS * 8u + so), "l"(bf.X + go) : "memory");
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s0 + 3u * MS * 8u + so), "l"(bf.A_q + go) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    } else {
        for (int idx = tid; idx < MP * MP; idx += T) {
            const int i = idx / MP, j = idx - i * MP;
            const bool in = i < M && j < M;
            const size_t gi = mo + (size_t)i * M + j;
            Lm[i * LD + j] = in ? bf.L[gi] : 0.0;
            Li[i * LD + j] = in ? bf.Li[gi] : 0.0;
            X[i * LD + j] = in ? bf.X[gi] : 0.0;
            B3[i * LD + j] = in ? bf.A_q[gi] : 0.0;
        }
    }
    const double* zg = bf.Z + (size_t)dm.R * ds.moff + (size_t)r * M;
    for (int i = tid; i < MP; i += T) {
        const bool in = i < M;
        al[i] = in ? bf.alpha[vo + i] : 0.0;
        cv[i] = in ? bf.c[vo + i] : 0.0;
        ab[i] = in ? bf.abar_q[vo + i] + bf.abar_spk[vo + i] : 0.0;
        zs[i] = in ? zg[i] : 0.0;
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();
    // A_q is stored lower: mirror it
    for (int idx = tid; idx < MP * MP; idx += T) {
        const int i = idx / MP, j = idx - i * MP;
        if (j > i) B3[i * LD + j] = B3[j * LD + i];
    }
    // y = Li abar: four threads per row (T / 4 = MP rows), two partial sums each -- a single thread per row was a
    // chain of up to 32 dependent (LDS, LDS, DFMA) steps executed by one warp while the other three waited
    {
        const int i = tid >> 2, part = tid & 3;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int p = part; p < MP; p += 8) {              // Li is stored with zeros above the diagonal
            s0 = fma(Li[i * LD + p], ab[p], s0);
            s1 = fma(Li[i * LD + p + 4], ab[p + 4], s1);
        }
        double sy = s0 + s1;
        sy += __shfl_xor_sync(0xffffffffu, sy, 1);
        sy += __shfl_xor_sync(0xffffffffu, sy, 2);
        if (part == 0) yv[i] = sy;
    }
    __syncthreads();
    // Xbar = tril(2 A X) - X -> B4 (lower tiles; the strict upper part of the diagonal tiles is zeroed)
    ipb_product<MT, false, false>(B3, X, it, lane,
        [&](int i_t, int jt, int& lo, int& hi) { lo = 2 * jt; hi = jt <= i_t ? KS - 1 : -1; },
        [&](int i, int j, double v) { B4[i * LD + j] = (j <= i) ? 2.0 * v - X[i * LD + j] : 0.0; });
    // mbar = Li^T (y - c), four threads per column
    {
        const int j = tid >> 2, part = tid & 3;
        double s0 = 0.0, s1 = 0.0;
#pragma unroll
        for (int i = part; i < MP; i += 8) {
            s0 = fma(Li[i * LD + j], yv[i] - cv[i], s0);
            s1 = fma(Li[(i + 4) * LD + j], yv[i + 4] - cv[i + 4], s1);
        }
        double sm_ = s0 + s1;
        sm_ += __shfl_xor_sync(0xffffffffu, sm_, 1);
        sm_ += __shfl_xor_sync(0xffffffffu, sm_, 2);
        if (part == 0) mb[j] = sm_;
    }
    __syncthreads();
    // T = Li^T tril(Xbar) -> B5 (full)      T(i,j) = sum_{p >= max(i,j)} Li(p,i) Xbar(p,j)
    ipb_product<MT, true, false>(Li, B4, it, lane,
        [&](int i_t, int jt, int& lo, int& hi) { lo = 2 * max(i_t, jt); hi = KS - 1; },
        [&](int i, int j, double v) { B5[i * LD + j] = v; });
    __syncthreads();
    if (need_post) {
        double* gm = bf.gm + (size_t)dm.R * ds.moff + (size_t)r * M;
        for (int i = tid; i < M; i += T) gm[i] = mb[i];
        double* gcv = bf.gcholvec + (size_t)dm.R * ds.poff + (size_t)r * ds.P;
        const double* cvec = bf.cholvec + (size_t)dm.R * ds.poff + (size_t)r * ds.P;
        for (int p = tid; p < ds.P; p += T) {             // coalesced over the packed row-major tril vector
            int i = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
            while (i * (i + 1) / 2 > p) --i;
            while ((i + 1) * (i + 2) / 2 <= p) ++i;
            const int j = p - i * (i + 1) / 2;
            gcv[p] = B5[i * LD + j] + (i == j ? 1.0 / cvec[p] : 0.0);
        }
    }
    if (!need_kz) return;
    // B4 = X^T A      (i,j) = sum_{p >= i} X(p,i) A(p,j)
    ipb_product<MT, true, false>(X, B3, it, lane,
        [&](int i_t, int, int& lo, int& hi) { lo = 2 * i_t; hi = KS - 1; },
        [&](int i, int j, double v) { B4[i * LD + j] = v; });
    __syncthreads();
    // B3 = E2 = X (X^T A) - A   (each element of B3 is read and written by its own thread only)
    ipb_product<MT, false, false>(X, B4, it, lane,
        [&](int i_t, int, int& lo, int& hi) { lo = 0; hi = 2 * i_t + 1; },
        [&](int i, int j, double v) { B3[i * LD + j] = v - B3[i * LD + j]; });
    __syncthreads();
    // Lbar (lower) -> B4 = -2 Li^T E2 - alpha y^T - mbar c^T - diag(1/L_ii) [- T X^T below]
    ipb_product<MT, true, false>(Li, B3, it, lane,
        [&](int i_t, int jt, int& lo, int& hi) { lo = 2 * i_t; hi = jt <= i_t ? KS - 1 : -1; },
        [&](int i, int j, double v) {
            double s = 0.0;
            if (j <= i && i < M) {
                s = -2.0 * v - al[i] * yv[j] - mb[i] * cv[j];
                if (i == j) s -= 1.0 / Lm[i * LD + i];
            }
            B4[i * LD + j] = s;
        });
    for (int idx = tid; idx < MP * MP; idx += T) {            // the skipped upper tiles of Lbar are zero
        const int i = idx / MP, j = idx - i * MP;
        if ((j >> 3) > (i >> 3)) B4[i * LD + j] = 0.0;
    }
    __syncthreads();
    //   (T X^T)(i,j) = sum_{p <= j} T(i,p) X(j,p)
    ipb_product<MT, false, true>(B5, X, it, lane,
        [&](int i_t, int jt, int& lo, int& hi) { lo = 0; hi = jt <= i_t ? 2 * jt + 1 : -1; },
        [&](int i, int j, double v) { if (j <= i) B4[i * LD + j] -= v; });
    __syncthreads();
    // P = Phi(L^T Lbar) -> B3 lower       (i,j) = sum_{p >= i} L(p,i) Lbar(p,j)      (diagonal: P_ii = s/2, S_ii = s)
    ipb_product<MT, true, false>(Lm, B4, it, lane,
        [&](int i_t, int jt, int& lo, int& hi) { lo = 2 * i_t; hi = jt <= i_t ? KS - 1 : -1; },
        [&](int i, int j, double v) { if (j <= i) B3[i * LD + j] = v; });
    __syncthreads();
    for (int idx = tid; idx < MP * MP; idx += T) {          // S = P + P^T
        const int i = idx / MP, j = idx - i * MP;
        if (j > i) B3[i * LD + j] = B3[j * LD + i];
    }
    __syncthreads();
    // U1 = S Li -> B4      (i,j) = sum_{p >= j} S(i,p) Li(p,j)
    ipb_product<MT, false, false>(B3, Li, it, lane,
        [&](int, int jt, int& lo, int& hi) { lo = 2 * jt; hi = KS - 1; },
        [&](int i, int j, double v) { B4[i * LD + j] = v; });
    __syncthreads();
    // Kbar = 0.5 Li^T U1 -> B5      (i,j) = sum_{p >= i} Li(p,i) U1(p,j)
    ipb_product<MT, true, false>(Li, B4, it, lane,
        [&](int i_t, int, int& lo, int& hi) { lo = 2 * i_t; hi = KS - 1; },
        [&](int i, int j, double v) { B5[i * LD + j] = 0.5 * v; });
    // the exp / sincos tables of the last phase live in B3, which is free now (1.5 KB of static shared memory would
    // cost the fourth resident CTA: 4 x 58 112 B is exactly the 227 KB of an SM)
    double* etab = B3;
    double2* sctab = reinterpret_cast<double2*>(B3 + 64);
    svgpfa_load_exp_tab64(etab);
    if (kc.type == SVGPFA_KERNEL_PERIODIC) svgpfa_load_sincos_tab<1>(sctab);
    __syncthreads();
    // dZ_i = 2 sum_j Kbar_ij dkappa/ddelta(z_i - z_j);  dtheta = sum_ij Kbar_ij dkappa/dtheta:
    // thread <-> (row i, quarter of the columns), two kernel evaluations per iteration (kappa_vals_n)
    double t0 = 0.0, t1 = 0.0;
    double* gZ = bf.gZ + (size_t)dm.R * ds.moff + (size_t)r * M;
    {
        const int i = tid >> 2, part = tid & 3;          // T / 4 = MP rows
        double dz = 0.0;
        const double zi = zs[i];
        constexpr int NE = (2 * MT) % 4 == 0 ? 4 : 2;     // MP / 4 = 2 MT columns per thread
#pragma unroll 1
        for (int j0 = part * (MP / 4); j0 < (part + 1) * (MP / 4); j0 += NE) {
            double dl[NE], kv[NE], qq[NE], s2x[NE];
#pragma unroll
            for (int e = 0; e < NE; ++e) dl[e] = zi - zs[j0 + e];
            kappa_vals_n<NE>(kc, dl, etab, sctab, kv, qq, s2x);
#pragma unroll
            for (int e = 0; e < NE; ++e) {
                const double h = (i < M && j0 + e < M) ? B5[i * LD + j0 + e] * kv[e] : 0.0;
                if (kc.type == SVGPFA_KERNEL_EXPQUAD) {
                    dz = fma(h, dl[e], dz);
                    t0 = fma(h, qq[e], t0);
                } else {
                    const double hs = h * s2x[e];
                    dz += hs;
                    t0 = fma(h, qq[e], t0);
                    t1 = fma(hs, dl[e], t1);
                }
            }
        }
        dz += __shfl_xor_sync(0xffffffffu, dz, 1);
        dz += __shfl_xor_sync(0xffffffffu, dz, 2);
        // dkappa/ddelta = kappa (delta | sin 2x) dd
        if (i < M && part == 0 && (flags & SVGPFA_GRAD_INDLOCS)) gZ[i] = 2.0 * kc.dd * dz + bf.dz_acc[vo + i];
    }
    if (flags & SVGPFA_GRAD_KERNEL) {
        const double s0 = block_sum(kc.dl * t0, red);
        const double s1 = block_sum(kc.dp * t1, red);
        if (tid == 0) {
            double* dth = bf.dth_part + (size_t)r * dm.TH + ds.thoff;
            dth[0] += s0;
            if (ds.nth > 1) dth[1] += s1;
        }
    }
}

template <int MT>
void launch_ipb_mma(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, cudaStream_t st) {
    constexpr int MP = 8 * MT, LD = MP + 4;
    const size_t smem = sizeof(double) * ((size_t)6 * MP * LD + 6 * MP);
    SVGPFA_ENSURE_SMEM(smem, indpoints_bwd_mma_kernel<MT>);
    indpoints_bwd_mma_kernel<MT><<<dim3(svgpfa_ntrials(dims), dims->K), 32 * MT, smem, st>>>(*dims, *buf, flags);
}

size_t ip_smem(int Mmax, int nmat, int nvec) {
    const int MP = (Mmax + 7) / 8 * 8, ld = MP + 4;      // covers both the odd-ld (M|1) and the padded (MP+4) layouts
    return sizeof(double) * ((size_t)nmat * MP * ld + (size_t)nvec * MP);
}

}  // namespace

extern "C" int svgpfa_kzz_chol_fwd(const svgpfa_dims* dims, const svgpfa_buffers* buf, void* stream) {
    if (!dims || !buf || dims->Mmax > SVGPFA_MAX_M || dims->Mmax < 1) return svgpfa_set_error(SVGPFA_E_ARG, "kzz_chol_fwd", cudaSuccess);
    if (dims->R == 0) return SVGPFA_OK;
    if (dims->Mmax <= 32) {
        const int nprob = svgpfa_ntrials(dims) * dims->K;
        const size_t wsm = sizeof(double) * KC_WARPS * KC_WSM;
        kzz_chol_warp_kernel<false><<<(nprob + KC_WARPS - 1) / KC_WARPS, 32 * KC_WARPS, wsm, (cudaStream_t)stream>>>(*dims, *buf, nprob);
    } else {
        kzz_chol_pair_kernel<false><<<dim3(svgpfa_ntrials(dims), dims->K), KP_N, 0, (cudaStream_t)stream>>>(*dims, *buf);
    }
    SVGPFA_CHECK_LAUNCH("kzz_chol_fwd");
    return SVGPFA_OK;
}

// svgpfa_kzz_chol_fwd + svgpfa_indpoints_fwd in one launch (M <= 32; the orchestrator uses it whenever both stages run).
// Returns false when the shape is outside this path.
bool svgpfa_try_chol_indpoints_fused(const svgpfa_dims* dims, const svgpfa_buffers* buf, cudaStream_t st) {
    if (dims->Mmax > 64) return false;
    const int nprob = svgpfa_ntrials(dims) * dims->K;
    if (nprob == 0) return true;
    if (dims->Mmax > 32) {                              // two warps per matrix
        kzz_chol_pair_kernel<true><<<dim3(svgpfa_ntrials(dims), dims->K), KP_N, 0, st>>>(*dims, *buf);
        return true;
    }
    const size_t wsm = sizeof(double) * KC_WARPS * KC_WSM;
    kzz_chol_warp_kernel<true><<<(nprob + KC_WARPS - 1) / KC_WARPS, 32 * KC_WARPS, wsm, st>>>(*dims, *buf, nprob);
    return true;
}

extern "C" int svgpfa_indpoints_fwd(const svgpfa_dims* dims, const svgpfa_buffers* buf, void* stream) {
    if (!dims || !buf || dims->Mmax > SVGPFA_MAX_M) return svgpfa_set_error(SVGPFA_E_ARG, "indpoints_fwd", cudaSuccess);
    if (dims->R == 0) return SVGPFA_OK;
    if (dims->Mmax <= 32) {
        const int nprob = svgpfa_ntrials(dims) * dims->K;
        const size_t wsm = sizeof(double) * KC_WARPS * KC_WSM;
        indpoints_fwd_warp_kernel<<<(nprob + KC_WARPS - 1) / KC_WARPS, 32 * KC_WARPS, wsm, (cudaStream_t)stream>>>(*dims, *buf, nprob);
    } else {
        const size_t smem = ip_smem(dims->Mmax, 2, 2);
        SVGPFA_ENSURE_SMEM(smem, indpoints_fwd_kernel);
        indpoints_fwd_kernel<<<dim3(svgpfa_ntrials(dims), dims->K), IP_THREADS, smem, (cudaStream_t)stream>>>(*dims, *buf);
    }
    SVGPFA_CHECK_LAUNCH("indpoints_fwd");
    return SVGPFA_OK;
}

extern "C" int svgpfa_indpoints_bwd(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, void* stream) {
    if (!dims || !buf || dims->Mmax > SVGPFA_MAX_M) return svgpfa_set_error(SVGPFA_E_ARG, "indpoints_bwd", cudaSuccess);
    if (dims->R == 0) return SVGPFA_OK;
    // compile-time shapes for every M <= 64 (MT = ceil(M / 8) warps per matrix; M > 32: one CTA per SM, 209 KB)
    cudaStream_t st = (cudaStream_t)stream;
    switch ((dims->Mmax + 7) / 8) {
        case 1: launch_ipb_mma<1>(dims, buf, flags, st); break;
        case 2: launch_ipb_mma<2>(dims, buf, flags, st); break;
        case 3: launch_ipb_mma<3>(dims, buf, flags, st); break;
        case 4: launch_ipb_mma<4>(dims, buf, flags, st); break;
        case 5: launch_ipb_mma<5>(dims, buf, flags, st); break;
        case 6: launch_ipb_mma<6>(dims, buf, flags, st); break;
        case 7: launch_ipb_mma<7>(dims, buf, flags, st); break;
        default: launch_ipb_mma<8>(dims, buf, flags, st); break;
    }
    SVGPFA_CHECK_LAUNCH("indpoints_bwd");
    return SVGPFA_OK;
}

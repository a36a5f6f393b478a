// Shared device helpers for the svGPFA lower-bound kernels (sm_100a, float64 throughout).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "svgpfa_b200.h"

#include "exp_table.cuh"

#define SVGPFA_PI 3.14159265358979323846

// exp(x) for x <= 0 -- every covariance-kernel value is exp of a non-positive number.
// On this part an FP64 instruction occupies the warp scheduler's issue port for 2 cycles (measured: cycles
// per iteration = 2*N_fp64 + N_other, profiles/README.md), so the routine minimises FP64 operations:
//   n = rint(x * 2048/ln2), r = x - n ln2/2048 (|r| <= 1.7e-4), exp(x) = 2^(n>>11) * T[n & 2047] * (1 + r + r^2/2 + r^3/6)
// with T = 2^(j/2048) in SHARED memory (16 KB per CTA): 7 FP64 instructions (libdevice exp: ~17 + branches).
// Relative error <= 3e-16 for x > -40; the single-FMA reduction adds <= 8e-17 |x| (absolute error <= 3e-17).
// Arguments below -707 return ~1e-307 instead of a subnormal (tests/test_gpu_kernels.py::test_exp_neg_accuracy).
__constant__ double svgpfa_expc[2] = {SVGPFA_EXP_INV_L, SVGPFA_EXP_L};

__device__ __forceinline__ void svgpfa_load_exp_tab(double* tab) {
    for (int i = threadIdx.x; i < SVGPFA_EXP_TAB_SIZE; i += blockDim.x) tab[i] = svgpfa_exp2_tab_g[i];
}

__device__ __forceinline__ double svgpfa_exp_neg(double x, const double* __restrict__ tab) {
    // clamp to > -707 on the integer pipe (x <= 0: a larger magnitude is a larger unsigned high word)
    const unsigned hi = min((unsigned)__double2hiint(x), 0xC08617FFu);
    x = __hiloint2double((int)hi, __double2loint(x));
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52: round-to-nearest-integer trick
    const double t = fma(x, svgpfa_expc[0], MAGIC);
    const double nd = t - MAGIC;
    const int n = __double2loint(t);
    const double r = fma(nd, -svgpfa_expc[1], x);
    double q = fma(r, 1.0 / 6.0, 0.5);
    q = fma(q, r, 1.0);
    const double T = tab[n & (SVGPFA_EXP_TAB_SIZE - 1)];
    const double e = fma(T * r, q, T);                        // T * exp(r), in [0.99, 2.01)
    return __hiloint2double(__double2hiint(e) + ((n >> SVGPFA_EXP_TAB_BITS) << 20), __double2loint(e));
}

// Small-table variant for kernels that cannot spare 16 KB of shared memory: 64 entries (every 32nd of the big
// table), degree-5 polynomial, 9 FP64 instructions.  Same clamp and error behaviour.
__device__ __forceinline__ void svgpfa_load_exp_tab64(double* tab) {
    for (int i = threadIdx.x; i < 64; i += blockDim.x) tab[i] = svgpfa_exp2_tab_g[32 * i];
}

__device__ __forceinline__ double svgpfa_exp_neg64(double x, const double* __restrict__ tab) {
    const unsigned hi = min((unsigned)__double2hiint(x), 0xC08617FFu);
    x = __hiloint2double((int)hi, __double2loint(x));
    const double MAGIC = 6755399441055744.0;
    const double t = fma(x, SVGPFA_EXP_INV_L * 0.03125, MAGIC);      // 64 / ln2
    const double nd = t - MAGIC;
    const int n = __double2loint(t);
    const double r = fma(nd, -32.0 * SVGPFA_EXP_L, x);               // ln2 / 64, |r| <= 5.5e-3
    double q = fma(r, 1.0 / 120.0, 1.0 / 24.0);
    q = fma(q, r, 1.0 / 6.0);
    q = fma(q, r, 0.5);
    q = fma(q, r, 1.0);
    const double T = tab[n & 63];
    const double e = fma(T * r, q, T);
    return __hiloint2double(__double2hiint(e) + ((n >> 6) << 20), __double2loint(e));
}

// Per-latent kernel constants derived from theta (kernels.py:33-46, 73-85).
//   expquad : kappa = s2 exp(nh d^2),            nh = -0.5 / l^2
//   periodic: kappa = s2 exp(nh sin^2(pi d/p)),  nh = -2 / l^2
// l = theta[0] / lengthscaleScale, p = theta[1] / periodScale.
struct KConst {
    int type;
    double s2;      // scale^2
    double nh;      // see above
    double invp;    // periodic: 1/p
    double dl;      // d kappa / d theta0 = kappa * q * dl,  q = d^2 (expquad) or sin^2 (periodic)
    double dp;      // periodic: d kappa / d theta1 = kappa * sin(2 pi d/p) * d * dp
    double dd;      // d kappa / d delta = kappa * (d or sin(2 pi d/p)) * dd
};

__device__ __forceinline__ KConst make_kconst(const svgpfa_latent_desc& ds, const double* __restrict__ theta,
                                              const double* __restrict__ kscale, int k) {
    KConst kc;
    kc.type = ds.ktype;
    kc.s2 = kscale[4 * k + 0];
    const double ils = kscale[4 * k + 1];
    const double l = theta[ds.thoff] * ils;
    if (ds.ktype == SVGPFA_KERNEL_EXPQUAD) {
        kc.nh = -0.5 / (l * l);
        kc.invp = 0.0;
        kc.dl = ils / (l * l * l);
        kc.dp = 0.0;
        kc.dd = -1.0 / (l * l);
    } else {
        const double ips = kscale[4 * k + 2];
        const double p = theta[ds.thoff + 1] * ips;
        kc.nh = -2.0 / (l * l);
        kc.invp = 1.0 / p;
        kc.dl = 4.0 * ils / (l * l * l);
        kc.dp = ips * 2.0 * SVGPFA_PI / (l * l * p * p);
        kc.dd = -2.0 * SVGPFA_PI / (p * l * l);
    }
    return kc;
}

// kappa(delta) only.
__device__ __forceinline__ double kappa_val(const KConst& kc, double delta) {
    double q;
    if (kc.type == SVGPFA_KERNEL_EXPQUAD) {
        q = delta * delta;
    } else {
        const double s = sinpi(delta * kc.invp);
        q = s * s;
    }
    return kc.s2 * exp(kc.nh * q);
}

// kappa and its partial derivatives w.r.t. delta (= x - z), theta0 and theta1.
__device__ __forceinline__ void kappa_grad(const KConst& kc, double delta, double& kv, double& dk_dd,
                                           double& dk_dt0, double& dk_dt1) {
    if (kc.type == SVGPFA_KERNEL_EXPQUAD) {
        const double q = delta * delta;
        kv = kc.s2 * exp(kc.nh * q);
        dk_dd = kv * delta * kc.dd;
        dk_dt0 = kv * q * kc.dl;
        dk_dt1 = 0.0;
    } else {
        double s, c;
        sincospi(delta * kc.invp, &s, &c);
        const double q = s * s;
        const double s2x = 2.0 * s * c;
        kv = kc.s2 * exp(kc.nh * q);
        dk_dd = kv * s2x * kc.dd;
        dk_dt0 = kv * q * kc.dl;
        dk_dt1 = kv * s2x * delta * kc.dp;
    }
}

// kappa with the 2048-entry table exp (tab in shared memory, svgpfa_load_exp_tab).
__device__ __forceinline__ double kappa_val_big(const KConst& kc, double delta, const double* __restrict__ tab) {
    double q;
    if (kc.type == SVGPFA_KERNEL_EXPQUAD) {
        q = delta * delta;
    } else {
        const double s = sinpi(delta * kc.invp);
        q = s * s;
    }
    return kc.s2 * svgpfa_exp_neg(kc.nh * q, tab);
}

// Same as kappa_val / kappa_grad with the 64-entry table exp (tab in shared memory).
__device__ __forceinline__ double kappa_val_t(const KConst& kc, double delta, const double* __restrict__ tab) {
    double q;
    if (kc.type == SVGPFA_KERNEL_EXPQUAD) {
        q = delta * delta;
    } else {
        const double s = sinpi(delta * kc.invp);
        q = s * s;
    }
    return kc.s2 * svgpfa_exp_neg64(kc.nh * q, tab);
}

__device__ __forceinline__ void kappa_grad_t(const KConst& kc, double delta, const double* __restrict__ tab, double& kv,
                                             double& dk_dd, double& dk_dt0, double& dk_dt1) {
    if (kc.type == SVGPFA_KERNEL_EXPQUAD) {
        const double q = delta * delta;
        kv = kc.s2 * svgpfa_exp_neg64(kc.nh * q, tab);
        dk_dd = kv * delta * kc.dd;
        dk_dt0 = kv * q * kc.dl;
        dk_dt1 = 0.0;
    } else {
        double s, c;
        sincospi(delta * kc.invp, &s, &c);
        const double q = s * s;
        const double s2x = 2.0 * s * c;
        kv = kc.s2 * svgpfa_exp_neg64(kc.nh * q, tab);
        dk_dd = kv * s2x * kc.dd;
        dk_dt0 = kv * q * kc.dl;
        dk_dt1 = kv * s2x * delta * kc.dp;
    }
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum; result valid in thread 0.  `red` = shared scratch of >= 32 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        for (int i = 0; i < nw; ++i) t += red[i];
    }
    return t;
}

__host__ __device__ __forceinline__ int round_up(int x, int m) { return (x + m - 1) / m * m; }

// host-side error plumbing (api.cu)
int svgpfa_set_error(int code, const char* where, cudaError_t ce);
#define SVGPFA_CHECK_LAUNCH(where)                                          \
    do {                                                                    \
        cudaError_t _e = cudaGetLastError();                                \
        if (_e != cudaSuccess) return svgpfa_set_error(SVGPFA_E_CUDA, where, _e); \
    } while (0)

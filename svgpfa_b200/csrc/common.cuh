// Shared device helpers for the svGPFA lower-bound kernels (sm_100a, float64 throughout).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include <atomic>

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

// 2^(-w2 / 256) for 0 <= w2 <= SVGPFA_EXP2M_LIMIT, for callers that work in pre-scaled coordinates (argument already
// in units of 1/256 octave, e.g. w2 = d^2 * 256 / (2 ln2 l^2)); no argument scaling and no clamp per evaluation.
//   n = rint(-w2), u = w2 + n in [-1/2, 1/2] (exact), 2^(-w2/256) = 2^(n>>8) T[n & 255] exp(-u ln2/256), degree 4.
// The 2048-entry table of svgpfa_exp_neg costs ~6 shared-memory wavefronts per warp lookup (random 8-byte gathers:
// ncu showed the LSU data pipe 80 % busy in the spike kernel).  Here the table has 256 entries, each REPLICATED 16
// times (32 KB): lane l reads replica l & 15, so the 16 lanes of a half-warp always hit 16 different bank pairs --
// 2 wavefronts, the minimum for a 64-bit load -- at the price of one more polynomial term: 8 FP64 + 4 other
// instructions.  The copy in shared memory has j << 12 subtracted from the high word of entry j, so that ONE
// integer add of n << 12 both restores the entry and applies the binary exponent n >> 8.
// Relative error <= 3.5e-16 (host check against long-double exp2; tests/test_gpu_kernels.py::test_exp2m_accuracy).
// Above the limit the exponent field would wrap: callers clamp (svgpfa_exp2m_clamp) or prove the bound.
#define SVGPFA_EXP2M_BITS 8
#define SVGPFA_EXP2M_REP 16
#define SVGPFA_EXP2M_TAB_BYTES (8 * SVGPFA_EXP2M_REP << SVGPFA_EXP2M_BITS)        /* 32 KB */
#define SVGPFA_EXP2M_INV_L (SVGPFA_EXP_INV_L * 0.125)                              /* 256 / ln2 */
#define SVGPFA_EXP2M_L (SVGPFA_EXP_L * 8.0)                                        /* ln2 / 256 */
#define SVGPFA_EXP2M_LIMIT 2.61e5            /* (n >> 8) >= -1020: the result stays a normal number */

// tab: SVGPFA_EXP2M_TAB_BYTES of shared memory, [entry][replica]
__device__ __forceinline__ void svgpfa_load_exp2m_tab(double* tab) {
    for (int i = threadIdx.x; i < (SVGPFA_EXP2M_REP << SVGPFA_EXP2M_BITS); i += blockDim.x) {
        const int j = i / SVGPFA_EXP2M_REP;
        const double v = svgpfa_exp2_tab_g[j << (SVGPFA_EXP_TAB_BITS - SVGPFA_EXP2M_BITS)];
        tab[i] = __hiloint2double(__double2hiint(v) - (j << (20 - SVGPFA_EXP2M_BITS)), __double2loint(v));
    }
}

// 32-bit shared-window address of this lane's replica of entry 0
__device__ __forceinline__ unsigned svgpfa_exp2m_lane_tab(const double* tab) {
    return (unsigned)__cvta_generic_to_shared(tab) + 8u * (threadIdx.x & (SVGPFA_EXP2M_REP - 1));
}

__device__ __forceinline__ double svgpfa_exp2m(double w2, unsigned lane_tab) {
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52
    const double L = SVGPFA_EXP2M_L;
    const double C1 = L, C2 = L * L / 2.0, C3 = L * L * L / 6.0, C4 = L * L * L * L / 24.0;
    const double t = MAGIC - w2;
    const double nd = t - MAGIC;
    const double u = w2 + nd;
    const int n = __double2loint(t);
    double q = fma(u, C4, -C3);
    q = fma(u, q, C2);
    q = fma(-u, q, C1);                                       // exp(-uL) = 1 - u q
    // entry address = lane_tab + (n & 255) * 128 (8 B * 16 replicas): one LOP3 and one IMAD
    double Tb;
    asm("{\n\t.reg .u32 i, a;\n\tand.b32 i, %1, 255;\n\tmad.lo.u32 a, i, 128, %2;\n\tld.shared.f64 %0, [a];\n\t}"
        : "=d"(Tb) : "r"(n), "r"(lane_tab));
    const double T = __hiloint2double(__double2hiint(Tb) + (n << (20 - SVGPFA_EXP2M_BITS)), __double2loint(Tb));
    // T * (1 - u q): two instructions with two register operands each (a DFMA with three distinct register operands
    // costs ~3.5 issue cycles on this part against 2, tools/probe_issue.py)
    return T * fma(-u, q, 1.0);
}

// NE independent evaluations of svgpfa_exp2m written stage by stage.  The opaque asm statements between the stages
// keep the compiler from serialising the NE dependency chains (it otherwise emits one chain after the other to save
// registers, and the FP64 latency of ~10 cycles per dependent instruction is then exposed; ncu: 43 % "wait" stalls).
#define SVGPFA_PIN4(x) asm volatile("" : "+d"(x[0]), "+d"(x[1]), "+d"(x[2]), "+d"(x[3]))
template <int NE>
__device__ __forceinline__ void svgpfa_pin(double (&x)[NE]) {
    if constexpr (NE == 8) asm volatile("" : "+d"(x[0]), "+d"(x[1]), "+d"(x[2]), "+d"(x[3]), "+d"(x[4]), "+d"(x[5]), "+d"(x[6]), "+d"(x[7]));
    else if constexpr (NE == 4) asm volatile("" : "+d"(x[0]), "+d"(x[1]), "+d"(x[2]), "+d"(x[3]));
    else if constexpr (NE == 2) asm volatile("" : "+d"(x[0]), "+d"(x[1]));
    else asm volatile("" : "+d"(x[0]));
}

// VAR (bit mask, measured with tools/probe_eval.py):
//   1  degree-3 polynomial instead of degree 4: the Taylor series of exp(-y), |y| <= h = ln2/512, truncated after y^4
//      with y^4 replaced by its best quadratic on [-h, h] (Chebyshev economisation: y^4 ~ h^2 y^2 - h^4/8), i.e.
//      exp(-y) ~ (1 - h^4/192) - y + (1/2 + h^2/24) y^2 - y^3/6, maximum relative error 1.8e-14 -- four orders inside
//      BASELINE.json's 1e-10 -- and one FP64 instruction fewer per evaluation;
//   2  the rounded integer comes back through the conversion unit (I2F.F64.S32 of the low word of MAGIC - w2)
//      instead of a second FP64 subtraction: one more FP64 issue slot moved off the FP64 pipe.
template <int NE, int VAR = 0>
__device__ __forceinline__ void svgpfa_exp2m_n(const double (&w2)[NE], unsigned lane_tab, double (&out)[NE]) {
    const double MAGIC = 6755399441055744.0;
    const double L = SVGPFA_EXP2M_L, H = 0.5 * SVGPFA_EXP2M_L;
    const double C1 = L, C2 = L * L / 2.0, C3 = L * L * L / 6.0, C4 = L * L * L * L / 24.0;
    const double E0 = 1.0 - H * H * H * H / 192.0, E2 = L * L * (0.5 + H * H / 24.0);
    double t[NE], u[NE], q[NE], T[NE];
    int n[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) t[e] = MAGIC - w2[e];
    svgpfa_pin(t);
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        n[e] = __double2loint(t[e]);
        double Tb;
        asm("{\n\t.reg .u32 i, a;\n\tand.b32 i, %1, 255;\n\tmad.lo.u32 a, i, 128, %2;\n\tld.shared.f64 %0, [a];\n\t}"
            : "=d"(Tb) : "r"(n[e]), "r"(lane_tab));
        T[e] = __hiloint2double(__double2hiint(Tb) + (n[e] << (20 - SVGPFA_EXP2M_BITS)), __double2loint(Tb));
    }
#pragma unroll
    for (int e = 0; e < NE; ++e) u[e] = w2[e] + ((VAR & 2) ? __int2double_rn(n[e]) : (t[e] - MAGIC));
    svgpfa_pin(u);
    if (VAR & 1) {
#pragma unroll
        for (int e = 0; e < NE; ++e) q[e] = fma(-u[e], C3, E2);
        svgpfa_pin(q);
#pragma unroll
        for (int e = 0; e < NE; ++e) q[e] = fma(-u[e], q[e], C1);
        svgpfa_pin(q);
#pragma unroll
        for (int e = 0; e < NE; ++e) q[e] = fma(-u[e], q[e], E0);
        svgpfa_pin(q);
    } else {
#pragma unroll
        for (int e = 0; e < NE; ++e) q[e] = fma(u[e], C4, -C3);
        svgpfa_pin(q);
#pragma unroll
        for (int e = 0; e < NE; ++e) q[e] = fma(u[e], q[e], C2);
        svgpfa_pin(q);
#pragma unroll
        for (int e = 0; e < NE; ++e) q[e] = fma(-u[e], q[e], C1);
        svgpfa_pin(q);
#pragma unroll
        for (int e = 0; e < NE; ++e) q[e] = fma(-u[e], q[e], 1.0);
        svgpfa_pin(q);
    }
#pragma unroll
    for (int e = 0; e < NE; ++e) out[e] = T[e] * q[e];
    svgpfa_pin(out);
}

// integer-pipe clamp of a non-negative double to SVGPFA_EXP2M_LIMIT (high words of non-negative doubles order
// like the values)
__device__ __forceinline__ double svgpfa_exp2m_clamp(double w2) {
    const unsigned hi = min((unsigned)__double2hiint(w2), 0x410FDC3Fu);      // 0x410FDC40 00000000 = 2.61e5
    return __hiloint2double((int)hi, __double2loint(w2));
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

// --------------------------------------------------------------------------------------------------------------
// NE independent kernel evaluations written stage by stage (the quadrature kernels' element-wise phases).
// Why: with two resident warps per scheduler these phases are latency-bound -- ptxas emits the evaluations of an
// unrolled loop one dependency chain after the other, and a chain is ~14 dependent FP64 instructions of ~10 cycles
// (ncu: "wait" was the top stall of both quadrature kernels).  The opaque pins keep NE chains in flight.
// --------------------------------------------------------------------------------------------------------------
// exp(x) for x <= 0, 64-entry table in shared memory (svgpfa_load_exp_tab64), same arithmetic as svgpfa_exp_neg64
// ANYSIGN: arguments of either sign, clamped to [-707, 709] (exp(709) = 8.2e307; the intensity integrand of a model whose
// rates overflow float64 is meaningless either way -- the reference would report an infinite bound).
template <int NE, bool ANYSIGN = false>
__device__ __forceinline__ void svgpfa_exp_neg64_n(double (&x)[NE], const double* __restrict__ tab, double (&out)[NE]) {
    const double MAGIC = 6755399441055744.0;
    double t[NE], r[NE], q[NE], T[NE];
    int n[NE];
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        if (ANYSIGN) {
            x[e] = fmin(fmax(x[e], -707.0), 709.0);
        } else {
            const unsigned hi = min((unsigned)__double2hiint(x[e]), 0xC08617FFu);      // clamp to > -707
            x[e] = __hiloint2double((int)hi, __double2loint(x[e]));
        }
    }
#pragma unroll
    for (int e = 0; e < NE; ++e) t[e] = fma(x[e], SVGPFA_EXP_INV_L * 0.03125, MAGIC);   // 64 / ln2
    svgpfa_pin(t);
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        n[e] = __double2loint(t[e]);
        T[e] = tab[n[e] & 63];
    }
#pragma unroll
    for (int e = 0; e < NE; ++e) r[e] = fma(t[e] - MAGIC, -32.0 * SVGPFA_EXP_L, x[e]);  // |r| <= ln2 / 128
    svgpfa_pin(r);
#pragma unroll
    for (int e = 0; e < NE; ++e) q[e] = fma(r[e], 1.0 / 120.0, 1.0 / 24.0);
    svgpfa_pin(q);
#pragma unroll
    for (int e = 0; e < NE; ++e) q[e] = fma(q[e], r[e], 1.0 / 6.0);
    svgpfa_pin(q);
#pragma unroll
    for (int e = 0; e < NE; ++e) q[e] = fma(q[e], r[e], 0.5);
    svgpfa_pin(q);
#pragma unroll
    for (int e = 0; e < NE; ++e) q[e] = fma(q[e], r[e], 1.0);
    svgpfa_pin(q);
#pragma unroll
    for (int e = 0; e < NE; ++e) {
        const double v = fma(T[e] * r[e], q[e], T[e]);
        out[e] = __hiloint2double(__double2hiint(v) + ((n[e] >> 6) << 20), __double2loint(v));
    }
}

// sin and cos of 2 pi x by table (x64 = 64 x): n = rint(64 x), u = 64 x - n in [-1/2, 1/2], theta = 2 pi u / 64,
//   sin(2 pi x) = S_n cos(theta) + C_n sin(theta),  cos(2 pi x) = C_n cos(theta) - S_n sin(theta)
// with (S_n, C_n) = sincos(2 pi n / 64) in shared memory and Taylor polynomials of degree 7 / 8 in theta (|theta| <=
// 0.0491: truncation < 5e-18).  19 FP64 instructions against ~40 (+ ~20 others and branches) of libdevice's sincospi.
// `tab` points at this lane's replica of entry 0; consecutive entries are STRIDE double2 apart (the spike kernel
// replicates the table 16 times to keep the gathers conflict-free; the quadrature kernels use one copy).
#define SVGPFA_SC_ENTRIES 64

template <int STRIDE>
__device__ __forceinline__ void svgpfa_load_sincos_tab(double2* tab) {
    for (int i = threadIdx.x; i < SVGPFA_SC_ENTRIES * STRIDE; i += blockDim.x) {
        double sv, cv;
        sincospi((double)(i / STRIDE) * (2.0 / SVGPFA_SC_ENTRIES), &sv, &cv);
        tab[i] = make_double2(sv, cv);
    }
}

template <int STRIDE>
__device__ __forceinline__ void svgpfa_sincos2pi_tab(double x64, const double2* __restrict__ tab, double& sv, double& cv) {
    const double MAGIC = 6755399441055744.0;                 // 1.5 * 2^52
    const double t = x64 + MAGIC;
    const double u = x64 - (t - MAGIC);
    const int n = __double2loint(t) & (SVGPFA_SC_ENTRIES - 1);
    const double th = u * (2.0 * SVGPFA_PI / SVGPFA_SC_ENTRIES), th2 = th * th;
    double ps = fma(th2, -1.0 / 5040.0, 1.0 / 120.0);
    ps = fma(th2, ps, -1.0 / 6.0);
    const double st = fma(th * th2, ps, th);                  // sin(theta)
    double pc = fma(th2, 1.0 / 40320.0, -1.0 / 720.0);
    pc = fma(th2, pc, 1.0 / 24.0);
    pc = fma(th2, pc, -0.5);
    const double ct = fma(th2, pc, 1.0);                      // cos(theta)
    const double2 sc = tab[n * STRIDE];
    sv = fma(sc.x, ct, sc.y * st);
    cv = fma(sc.y, ct, -sc.x * st);
}

// kappa(delta_e), e < NE, for one latent.  etab: 64-entry exp table, sctab: single-copy sincos table (periodic only).
// Periodic: sin^2(pi d/p) = (1 - cos(2 pi d/p)) / 2; q[e] returns that square (expquad: delta^2) and s2x[e] the
// sine of the doubled angle (expquad: unused) for the derivative formulas of kappa_grad.
template <int NE>
__device__ __forceinline__ void kappa_vals_n(const KConst& kc, const double (&delta)[NE], const double* __restrict__ etab,
                                             const double2* __restrict__ sctab, double (&kv)[NE], double (&q)[NE],
                                             double (&s2x)[NE]) {
    double x[NE];
    if (kc.type == SVGPFA_KERNEL_EXPQUAD) {
#pragma unroll
        for (int e = 0; e < NE; ++e) q[e] = delta[e] * delta[e];
    } else {
        const double invp64 = kc.invp * SVGPFA_SC_ENTRIES;
#pragma unroll
        for (int e = 0; e < NE; ++e) {
            double c2x;
            svgpfa_sincos2pi_tab<1>(delta[e] * invp64, sctab, s2x[e], c2x);
            q[e] = fmax(0.5 * (1.0 - c2x), 0.0);
        }
    }
#pragma unroll
    for (int e = 0; e < NE; ++e) x[e] = kc.nh * q[e];
    svgpfa_exp_neg64_n<NE>(x, etab, kv);
#pragma unroll
    for (int e = 0; e < NE; ++e) kv[e] *= kc.s2;
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

inline int svgpfa_ntrials(const svgpfa_dims* d) { return SVGPFA_NTRIALS(d); }

// Launch-path housekeeping done ONCE per (call site / template instantiation, device) instead of on every launch
// (function attributes are per device and context; cudaFuncSetAttribute and the occupancy query cost several
// microseconds each and used to sit in front of every kernel of every evaluation).  Idempotent bodies only: two host
// threads racing on the first launch both run the body.
#define SVGPFA_ONCE_PER_DEVICE(...)                                                     \
    do {                                                                                \
        static std::atomic<uint64_t> done_{0};                                          \
        int dev_ = 0;                                                                   \
        cudaGetDevice(&dev_);                                                           \
        const uint64_t bit_ = 1ull << (dev_ & 63);                                      \
        if (!(done_.load(std::memory_order_acquire) & bit_)) {                          \
            __VA_ARGS__;                                                                \
            done_.fetch_or(bit_, std::memory_order_release);                            \
        }                                                                               \
    } while (0)

// Dynamic shared-memory opt-in of a kernel, raised only when a launch needs more than any earlier one on this device.
// Usage: SVGPFA_ENSURE_SMEM(bytes, kernel<template, args>);
#define SVGPFA_ENSURE_SMEM(bytes, ...)                                                                          \
    do {                                                                                                        \
        static std::atomic<size_t> cur_[64];                                                                    \
        int dev_ = 0;                                                                                           \
        cudaGetDevice(&dev_);                                                                                   \
        if (cur_[dev_ & 63].load(std::memory_order_acquire) < (size_t)(bytes)) {                                \
            cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes));       \
            cudaFuncSetAttribute(__VA_ARGS__, cudaFuncAttributePreferredSharedMemoryCarveout,                   \
                                 cudaSharedmemCarveoutMaxShared);                                               \
            cur_[dev_ & 63].store((size_t)(bytes), std::memory_order_release);                                  \
        }                                                                                                       \
    } while (0)

// multiprocessor count of the current device, cached per device (api.cu)
int svgpfa_sm_count();

// host-side error plumbing (api.cu)
int svgpfa_set_error(int code, const char* where, cudaError_t ce);
#define SVGPFA_CHECK_LAUNCH(where)                                          \
    do {                                                                    \
        cudaError_t _e = cudaGetLastError();                                \
        if (_e != cudaSuccess) return svgpfa_set_error(SVGPFA_E_CUDA, where, _e); \
    } while (0)

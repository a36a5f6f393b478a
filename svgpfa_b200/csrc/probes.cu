// Measurement probes and test hooks -- NOT part of the product library (built into libsvgpfa_b200_probes.so,
// loaded only by bench.py's roofline leg, tools/probe_*.py and tests/test_gpu_kernels.py).
// See include/svgpfa_b200_probes.h.
#include <stdio.h>

#include "common.cuh"
#include "svgpfa_b200_probes.h"

namespace {

thread_local char g_perr[256] = "";

int probe_error(const char* where, cudaError_t ce) {
    snprintf(g_perr, sizeof(g_perr), "%s: %s", where, ce == cudaSuccess ? "bad argument" : cudaGetErrorString(ce));
    return ce == cudaSuccess ? SVGPFA_E_ARG : SVGPFA_E_CUDA;
}

// ------------------------------------------------------------------------------------------
// FP64 pipe probes
// ------------------------------------------------------------------------------------------
// Are DFMA (FP64 pipe) and mma.m8n8k4.f64 (tensor DMMA sub-pipe) independent?  MIX = 8 DFMA + NM DMMA per iteration.
template <int NF, int NM>
__global__ void __launch_bounds__(256) mix_probe_kernel(long iters, double* out) {
    const double av = 1.0 + 1e-9 * threadIdx.x, bv = 1.0 - 1e-9 * threadIdx.x;
    double c[8][2], a[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { c[e][0] = c[e][1] = 1e-3 * e; a[e] = 1e-3 * (threadIdx.x + 1) + 0.01 * e; }
    const double m = 0.999999, k = 1e-9;
    for (long i = 0; i < iters; ++i) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            if (e < NM)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[e][0]), "+d"(c[e][1]) : "d"(av), "d"(bv));
            if (e < NF) a[e] = fma(a[e], m, k);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int e = 0; e < 8; ++e) s += c[e][0] + c[e][1] + a[e];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// Issue-model probes: 8 independent FP64 chains per thread with ALL operands in distinct registers (the DFMA probe
// above has two constant operands), optionally interleaved with NI integer instructions or one shared-memory load
// per FP64 instruction.  OP: 0 = DFMA (3 register operands), 1 = DADD, 2 = DMUL.
template <int OP, int NI, int LDS>
__global__ void __launch_bounds__(256) issue_probe_kernel(long iters, double* out) {
    __shared__ double sm[256];
    sm[threadIdx.x] = 1e-9 * threadIdx.x;
    __syncthreads();
    double a[8], b[8], c[8];
    int k[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        a[e] = out[(threadIdx.x + e) & 255] * 1e-300 + 1e-3 * (threadIdx.x + 1) + 0.01 * e;
        b[e] = 0.999999 - 1e-9 * e + out[(threadIdx.x + 8 + e) & 255] * 1e-300;
        c[e] = 1e-9 * (e + 1) + out[(threadIdx.x + 16 + e) & 255] * 1e-300;
        k[e] = threadIdx.x + e;
    }
    for (long i = 0; i < iters; ++i) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            if (OP == 0) a[e] = fma(a[e], b[e], c[e]);
            else if (OP == 1) a[e] = a[e] + c[e];
            else a[e] = a[e] * b[e];
            if (NI >= 1) k[e] = k[e] * 3 + 1;
            if (NI >= 2) k[e] = (k[e] >> 3) ^ k[e];
            if (NI >= 3) k[e] = k[e] * 5 + 7;
            if (LDS) c[e] += sm[(k[e] + (int)i) & 255] * 0.0;   // 1 LDS (+1 DFMA) per step
        }
    }
    double s = 0.0;
#pragma unroll
    for (int e = 0; e < 8; ++e) s += a[e] + (double)k[e] + c[e];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// The spike kernel's evaluation sequence in isolation (no segments, no staging): per step 4 evaluations of
// w = t sc + zs, kappa = 2^(-w^2/256), pn += kappa, p1 += kappa w, p2 += kappa w^2 with t read from shared memory.
// MODE 0: as in the round-1 kernel; 1: without the moment accumulations (KGRAD = false); 2: table entry replaced by a
// constant (no table LDS); 3: without the spike-time LDS; 4: degree-3 polynomial; 5: I2F range reduction; 6: both;
// 7: both + pre-scaled spike times (w = ts + zs: a two-register DADD instead of a three-register DFMA).
template <int MODE>
__global__ void __maxnreg__(128) eval_probe_kernel(long iters, double* out) {
    __shared__ double tab[SVGPFA_EXP2M_TAB_BYTES / 8];
    __shared__ double ts[1024];
    svgpfa_load_exp2m_tab(tab);
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) ts[i] = 1e-3 * i;
    __syncthreads();
    const unsigned lane_tab = svgpfa_exp2m_lane_tab(tab);
    const double sc = 13.0 + out[threadIdx.x] * 1e-300, zs = -0.37 * (threadIdx.x & 31) - 1e-3 * (threadIdx.x >> 5);
    double pn = 0.0, p1 = 0.0, p2 = 0.0;
    for (long i = 0; i < iters; ++i) {
        const double* tp = ts + ((i * 4) & 1020);
        double t[4], w[4], w2[4], kv[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) t[e] = MODE == 3 ? 1e-3 * e + pn * 1e-300 : tp[e];
#pragma unroll
        for (int e = 0; e < 4; ++e) { w[e] = MODE == 7 ? t[e] + zs : fma(t[e], sc, zs); w2[e] = w[e] * w[e]; }
        if (MODE == 2) {
            const double MAGIC = 6755399441055744.0, L = SVGPFA_EXP2M_L;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const double tt = MAGIC - w2[e];
                const double u = w2[e] + (tt - MAGIC);
                const int n = __double2loint(tt);
                double q = fma(u, L * L * L * L / 24.0, -L * L * L / 6.0);
                q = fma(u, q, L * L / 2.0);
                q = fma(-u, q, L);
                const double T = __hiloint2double(0x3ff00000 + (n << 12), n & 255);
                kv[e] = fma(-(T * u), q, T);
            }
        } else if (MODE == 4) {
            svgpfa_exp2m_n<4, 1>(w2, lane_tab, kv);
        } else if (MODE == 5) {
            svgpfa_exp2m_n<4, 2>(w2, lane_tab, kv);
        } else if (MODE == 6 || MODE == 7) {
            svgpfa_exp2m_n<4, 3>(w2, lane_tab, kv);
        } else {
            svgpfa_exp2m_n<4>(w2, lane_tab, kv);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            pn += kv[e];
            if (MODE != 1) { p1 = fma(kv[e], w[e], p1); p2 = fma(kv[e], w2[e], p2); }
        }
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = pn + p1 + p2;
}

template <int KIND>
__global__ void __launch_bounds__(256) peak_probe_kernel(long iters, double* out) {
    __shared__ double etab[SVGPFA_EXP_TAB_SIZE];
    svgpfa_load_exp_tab(etab);
    __syncthreads();
    const double seed = 1e-3 * (threadIdx.x + 1) + 1e-7 * blockIdx.x;
    double a[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) a[e] = seed + 0.01 * e;
    const double m = 0.999999, c = 1e-9;
    for (long i = 0; i < iters; ++i) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            if (KIND == 0) a[e] = fma(a[e], m, c);
            else if (KIND == 1) a[e] = exp(-a[e] * 0.5) + c;           // 1 mul + 1 add + exp
            else if (KIND == 2) { double s, cs; sincospi(a[e], &s, &cs); a[e] = s * cs + 0.25; }
            else a[e] = svgpfa_exp_neg(-a[e] * 0.5, etab) + c;
        }
    }
    double s = 0.0;
#pragma unroll
    for (int e = 0; e < 8; ++e) s += a[e];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FP64 tensor-core probe: 8 independent mma.m8n8k4.f64 accumulator chains per warp (256 FMAs per instruction).
__global__ void __launch_bounds__(256) dmma_probe_kernel(long iters, double* out) {
    const double av = 1.0 + 1e-9 * threadIdx.x, bv = 1.0 - 1e-9 * threadIdx.x;
    double c[8][2];
#pragma unroll
    for (int e = 0; e < 8; ++e) c[e][0] = c[e][1] = 1e-3 * e;
    for (long i = 0; i < iters; ++i) {
#pragma unroll
        for (int e = 0; e < 8; ++e)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[e][0]), "+d"(c[e][1]) : "d"(av), "d"(bv));
    }
    double s = 0.0;
#pragma unroll
    for (int e = 0; e < 8; ++e) s += c[e][0] + c[e][1];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void exp_eval_kernel(const double* x, double* yf, double* yr, long n) {
    __shared__ double etab[SVGPFA_EXP_TAB_SIZE];
    svgpfa_load_exp_tab(etab);
    __syncthreads();
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        yf[i] = svgpfa_exp_neg(x[i], etab);
        yr[i] = exp(x[i]);
    }
}

__global__ void exp2m_eval_kernel(const double* w2, double* y, long n, int variant) {
    __shared__ double tab[SVGPFA_EXP2M_TAB_BYTES / 8];
    svgpfa_load_exp2m_tab(tab);
    __syncthreads();
    const unsigned lane_tab = svgpfa_exp2m_lane_tab(tab);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        if (variant == 3) {                       // what spike_tile_kernel evaluates: degree 3 + I2F (svgpfa_exp2m_n<., 3>)
            double a[1] = {svgpfa_exp2m_clamp(w2[i])}, o[1];
            svgpfa_exp2m_n<1, 3>(a, lane_tab, o);
            y[i] = o[0];
        } else {
            y[i] = svgpfa_exp2m(svgpfa_exp2m_clamp(w2[i]), lane_tab);
        }
    }
}

}  // namespace

extern "C" const char* svgpfa_probes_last_error(void) { return g_perr; }

extern "C" int svgpfa_peak_probe(int32_t kind, int32_t blocks, int64_t iters, double* out, void* stream) {
    if (!out || blocks < 1 || iters < 0) return probe_error("peak_probe", cudaSuccess);
    cudaStream_t st = (cudaStream_t)stream;
    switch (kind) {
        case 0: peak_probe_kernel<0><<<blocks, 256, 0, st>>>((long)iters, out); break;
        case 1: peak_probe_kernel<1><<<blocks, 256, 0, st>>>((long)iters, out); break;
        case 2: peak_probe_kernel<2><<<blocks, 256, 0, st>>>((long)iters, out); break;
        case 3: peak_probe_kernel<3><<<blocks, 256, 0, st>>>((long)iters, out); break;
        case 4: dmma_probe_kernel<<<blocks, 256, 0, st>>>((long)iters, out); break;
        case 5: mix_probe_kernel<8, 1><<<blocks, 256, 0, st>>>((long)iters, out); break;   // 8 DFMA + 1 DMMA
        case 6: mix_probe_kernel<8, 2><<<blocks, 256, 0, st>>>((long)iters, out); break;   // 8 DFMA + 2 DMMA
        case 7: mix_probe_kernel<8, 0><<<blocks, 256, 0, st>>>((long)iters, out); break;   // 8 DFMA
        case 8: mix_probe_kernel<0, 2><<<blocks, 256, 0, st>>>((long)iters, out); break;   // 2 DMMA
        case 9: mix_probe_kernel<8, 4><<<blocks, 256, 0, st>>>((long)iters, out); break;   // 8 DFMA + 4 DMMA
        case 10: issue_probe_kernel<0, 0, 0><<<blocks, 256, 0, st>>>((long)iters, out); break;  // DFMA, 3 registers
        case 11: issue_probe_kernel<1, 0, 0><<<blocks, 256, 0, st>>>((long)iters, out); break;  // DADD
        case 12: issue_probe_kernel<2, 0, 0><<<blocks, 256, 0, st>>>((long)iters, out); break;  // DMUL
        case 13: issue_probe_kernel<0, 1, 0><<<blocks, 256, 0, st>>>((long)iters, out); break;  // DFMA + 1 int
        case 14: issue_probe_kernel<0, 2, 0><<<blocks, 256, 0, st>>>((long)iters, out); break;  // DFMA + 2 int
        case 15: issue_probe_kernel<0, 3, 0><<<blocks, 256, 0, st>>>((long)iters, out); break;  // DFMA + 3 int
        case 16: issue_probe_kernel<0, 1, 1><<<blocks, 256, 0, st>>>((long)iters, out); break;  // 2 DFMA + 1 int + 1 LDS (+addr)
        case 20: eval_probe_kernel<0><<<blocks, 128, 0, st>>>((long)iters, out); break;
        case 21: eval_probe_kernel<1><<<blocks, 128, 0, st>>>((long)iters, out); break;
        case 22: eval_probe_kernel<2><<<blocks, 128, 0, st>>>((long)iters, out); break;
        case 23: eval_probe_kernel<3><<<blocks, 128, 0, st>>>((long)iters, out); break;
        case 24: eval_probe_kernel<4><<<blocks, 128, 0, st>>>((long)iters, out); break;
        case 25: eval_probe_kernel<5><<<blocks, 128, 0, st>>>((long)iters, out); break;
        case 26: eval_probe_kernel<6><<<blocks, 128, 0, st>>>((long)iters, out); break;
        case 27: eval_probe_kernel<7><<<blocks, 128, 0, st>>>((long)iters, out); break;
        default: return probe_error("peak_probe kind", cudaSuccess);
    }
    { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return probe_error("peak_probe", e_); }
    return SVGPFA_OK;
}
extern "C" int svgpfa_exp_neg_eval(const double* x, double* y_fast, double* y_ref, int64_t n, void* stream) {
    if (!x || !y_fast || !y_ref || n < 0) return probe_error("exp_neg_eval", cudaSuccess);
    if (n == 0) return SVGPFA_OK;
    exp_eval_kernel<<<256, 256, 0, (cudaStream_t)stream>>>(x, y_fast, y_ref, (long)n);
    { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return probe_error("exp_neg_eval", e_); }
    return SVGPFA_OK;
}

extern "C" int svgpfa_exp2m_eval(const double* w2, double* y, int64_t n, int32_t variant, void* stream) {
    if (!w2 || !y || n < 0) return probe_error("exp2m_eval", cudaSuccess);
    if (n == 0) return SVGPFA_OK;
    exp2m_eval_kernel<<<256, 256, 0, (cudaStream_t)stream>>>(w2, y, (long)n, (int)variant);
    { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return probe_error("exp2m_eval", e_); }
    return SVGPFA_OK;
}


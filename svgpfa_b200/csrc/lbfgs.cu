// Vector primitives of the device-resident, shard-aware L-BFGS.
//
// The reference optimises every step of its ECM loop with torch.optim.LBFGS (stats/svEM.py:218-294).  At config #5
// the E-step vector (m, cholVecs of every trial) has 2.2e8 entries: torch's two-loop recursion walks the history with
// one dot product and one axpy per stored pair and step (about 10 h n-vectors of traffic per iteration for a history
// of h pairs, each result read back by the host), and under trial sharding every one of those dot products would
// need its own collective.  svgpfa_b200/lbfgs.py runs the recursion in COEFFICIENT space instead (the direction is a
// linear combination of the stored s_i, y_i and the gradient; the recursion only needs their Gram matrix), so that an
// iteration touches the history exactly twice, with the kernels below:
//
//   lb_multidot   the new rows of the Gram matrix: every stored vector against up to three probe vectors (the new s,
//                 the new y, the new gradient) in ONE pass over the history -- (nv + 3 nv / 8) n doubles of traffic;
//   lb_combine    d = sum_i coef_i v_i in one pass, with g.d and max|d| (the line search's first questions) folded
//                 into the same pass;
//   lb_update     s = t d, y = g - g_prev, g_prev = g      (one pass instead of four tensor operations)
//   lb_step       x = x0 + t d                             (the line search's trial point, no clone / restore)
//   lb_stats      a.b, max|a|, sum|a|, max|b|              (directional derivative and optimality test of a trial point)
//
// All of them are HBM-bound streaming kernels: 16-byte loads, grids sized to the SM count, per-block partial results
// combined in block order by a second small kernel, so results are run-to-run reproducible and, under sharding, the
// only exchange is one small all-reduce of the partial Gram rows per iteration.
#include <cstdint>

#include "common.cuh"

namespace {

constexpr int LB_THREADS = 256;
constexpr int LB_GROUP = 8;                                  // stored vectors per CTA of the multidot kernel
constexpr int LB_MAXV = SVGPFA_LBFGS_MAX_VECS;
constexpr int LB_MAXB = SVGPFA_LBFGS_MAX_BLOCKS;

struct LbPtrs { const double* v[LB_MAXV]; };
struct LbCoef { double c[LB_MAXV]; };

__device__ __forceinline__ double2 ld2(const double* p) { return *reinterpret_cast<const double2*>(p); }

// elements [lo, hi) of block b when n elements are dealt to nb blocks in even-sized spans
__device__ __forceinline__ void lb_span(size_t n, int nb, int b, size_t& lo, size_t& hi) {
    size_t span = (n + nb - 1) / nb;
    span = (span + 1) & ~(size_t)1;
    lo = (size_t)b * span;
    hi = lo + span;
    if (lo > n) lo = n;
    if (hi > n) hi = n;
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// part[(group * 8 + v) * NP + j][block] = v-th vector of the group . probe j over the block's span
template <int NP>
__global__ void __launch_bounds__(LB_THREADS) lb_multidot_kernel(LbPtrs vp, LbPtrs pp, size_t n, double* __restrict__ part) {
    __shared__ double red[LB_THREADS / 32][LB_GROUP * NP];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nb = gridDim.x;
    const int g0 = blockIdx.y * LB_GROUP;
    size_t lo, hi;
    lb_span(n, nb, blockIdx.x, lo, hi);
    double acc[LB_GROUP][NP];
#pragma unroll
    for (int v = 0; v < LB_GROUP; ++v)
#pragma unroll
        for (int j = 0; j < NP; ++j) acc[v][j] = 0.0;
    size_t i = lo + 2 * (size_t)tid;
    for (; i + 1 < hi; i += 2 * LB_THREADS) {
        double2 p[NP], x[LB_GROUP];
#pragma unroll
        for (int j = 0; j < NP; ++j) p[j] = ld2(pp.v[j] + i);
#pragma unroll
        for (int v = 0; v < LB_GROUP; ++v) x[v] = ld2(vp.v[g0 + v] + i);
#pragma unroll
        for (int v = 0; v < LB_GROUP; ++v)
#pragma unroll
            for (int j = 0; j < NP; ++j) acc[v][j] = fma(x[v].y, p[j].y, fma(x[v].x, p[j].x, acc[v][j]));
    }
    if (i < hi) {                                            // odd tail of the last span
#pragma unroll
        for (int v = 0; v < LB_GROUP; ++v)
#pragma unroll
            for (int j = 0; j < NP; ++j) acc[v][j] = fma(vp.v[g0 + v][i], pp.v[j][i], acc[v][j]);
    }
#pragma unroll
    for (int v = 0; v < LB_GROUP; ++v)
#pragma unroll
        for (int j = 0; j < NP; ++j) {
            const double s = warp_sum(acc[v][j]);
            if (lane == 0) red[warp][v * NP + j] = s;
        }
    __syncthreads();
    if (tid < LB_GROUP * NP) {
        double s = 0.0;
#pragma unroll
        for (int w = 0; w < LB_THREADS / 32; ++w) s += red[w][tid];
        part[((size_t)g0 * NP + tid) * nb + blockIdx.x] = s;
    }
}

// out[o] = sum (or max, bit o of max_mask) over the nb block partials of output o: one warp per output, lane-strided
// loads (one round of memory latency; a thread per output walking 592 partials took 30-50 us) and a butterfly in a
// fixed order, so the result does not depend on the launch
constexpr int LB_RED_WARPS = 4;
__global__ void __launch_bounds__(32 * LB_RED_WARPS) lb_reduce_kernel(const double* __restrict__ part, int nb, int nout, uint64_t max_mask,
                                                                     double* __restrict__ out) {
    const int o = blockIdx.x * LB_RED_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (o >= nout) return;
    const double* p = part + (size_t)o * nb;
    if ((max_mask >> (o & 63)) & 1) {
        double m = 0.0;
        for (int b = lane; b < nb; b += 32) m = fmax(m, p[b]);
        m = warp_max(m);
        if (lane == 0) out[o] = m;
    } else {
        double s = 0.0;
        for (int b = lane; b < nb; b += 32) s += p[b];
        s = warp_sum(s);
        if (lane == 0) out[o] = s;
    }
}

// d = sum_v coef_v vec_v;  part[0][block] = g . d,  part[1][block] = max |d|
template <bool ACC>
__global__ void __launch_bounds__(LB_THREADS) lb_combine_kernel(LbPtrs vp, LbCoef cf, int nv, double* __restrict__ d,
                                                                const double* __restrict__ g, size_t n, double* __restrict__ part) {
    __shared__ double red[2][LB_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nb = gridDim.x;
    size_t lo, hi;
    lb_span(n, nb, blockIdx.x, lo, hi);
    double gd = 0.0, mx = 0.0;
    size_t i = lo + 2 * (size_t)tid;
    for (; i + 1 < hi; i += 2 * LB_THREADS) {
        double2 s = ACC ? ld2(d + i) : make_double2(0.0, 0.0);
#pragma unroll 8
        for (int v = 0; v < nv; ++v) {
            const double2 x = ld2(vp.v[v] + i);
            s.x = fma(cf.c[v], x.x, s.x);
            s.y = fma(cf.c[v], x.y, s.y);
        }
        *reinterpret_cast<double2*>(d + i) = s;
        const double2 gg = ld2(g + i);
        gd = fma(gg.y, s.y, fma(gg.x, s.x, gd));
        mx = fmax(mx, fmax(fabs(s.x), fabs(s.y)));
    }
    if (i < hi) {
        double s = ACC ? d[i] : 0.0;
        for (int v = 0; v < nv; ++v) s = fma(cf.c[v], vp.v[v][i], s);
        d[i] = s;
        gd = fma(g[i], s, gd);
        mx = fmax(mx, fabs(s));
    }
    gd = warp_sum(gd);
    mx = warp_max(mx);
    if (lane == 0) { red[0][warp] = gd; red[1][warp] = mx; }
    __syncthreads();
    if (tid == 0) {
        double s = 0.0, m = 0.0;
#pragma unroll
        for (int w = 0; w < LB_THREADS / 32; ++w) { s += red[0][w]; m = fmax(m, red[1][w]); }
        part[blockIdx.x] = s;
        part[nb + blockIdx.x] = m;
    }
}

// part[0..3][block] = a.b, max|a|, sum|a|, max|b|   (b may be null: a.b = max|b| = 0)
__global__ void __launch_bounds__(LB_THREADS) lb_stats_kernel(const double* __restrict__ a, const double* __restrict__ b, size_t n,
                                                              double* __restrict__ part) {
    __shared__ double red[4][LB_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nb = gridDim.x;
    size_t lo, hi;
    lb_span(n, nb, blockIdx.x, lo, hi);
    double ab = 0.0, ma = 0.0, sa = 0.0, mb = 0.0;
    size_t i = lo + 2 * (size_t)tid;
    for (; i + 1 < hi; i += 2 * LB_THREADS) {
        const double2 x = ld2(a + i);
        ma = fmax(ma, fmax(fabs(x.x), fabs(x.y)));
        sa += fabs(x.x) + fabs(x.y);
        if (b) {
            const double2 y = ld2(b + i);
            ab = fma(x.y, y.y, fma(x.x, y.x, ab));
            mb = fmax(mb, fmax(fabs(y.x), fabs(y.y)));
        }
    }
    if (i < hi) {
        ma = fmax(ma, fabs(a[i]));
        sa += fabs(a[i]);
        if (b) { ab = fma(a[i], b[i], ab); mb = fmax(mb, fabs(b[i])); }
    }
    ab = warp_sum(ab); sa = warp_sum(sa); ma = warp_max(ma); mb = warp_max(mb);
    if (lane == 0) { red[0][warp] = ab; red[1][warp] = ma; red[2][warp] = sa; red[3][warp] = mb; }
    __syncthreads();
    if (tid == 0) {
        double s0 = 0.0, m1 = 0.0, s2 = 0.0, m3 = 0.0;
#pragma unroll
        for (int w = 0; w < LB_THREADS / 32; ++w) {
            s0 += red[0][w]; m1 = fmax(m1, red[1][w]); s2 += red[2][w]; m3 = fmax(m3, red[3][w]);
        }
        part[blockIdx.x] = s0;
        part[nb + blockIdx.x] = m1;
        part[2 * nb + blockIdx.x] = s2;
        part[3 * nb + blockIdx.x] = m3;
    }
}

// s = t d, y = g - gp, gp = g
__global__ void __launch_bounds__(LB_THREADS) lb_update_kernel(double* __restrict__ s, double* __restrict__ y, const double* __restrict__ d,
                                                               double t, const double* __restrict__ g, double* __restrict__ gp, size_t n) {
    const size_t stride = 2 * (size_t)gridDim.x * LB_THREADS;
    size_t i = 2 * ((size_t)blockIdx.x * LB_THREADS + threadIdx.x);
    for (; i + 1 < n; i += stride) {
        const double2 dd = ld2(d + i), gg = ld2(g + i), pp = ld2(gp + i);
        *reinterpret_cast<double2*>(s + i) = make_double2(t * dd.x, t * dd.y);
        *reinterpret_cast<double2*>(y + i) = make_double2(gg.x - pp.x, gg.y - pp.y);
        *reinterpret_cast<double2*>(gp + i) = gg;
    }
    if (i < n) { s[i] = t * d[i]; y[i] = g[i] - gp[i]; gp[i] = g[i]; }
}

// x = x0 + t d; VEC: every pointer 16-byte aligned
template <bool VEC>
__global__ void __launch_bounds__(LB_THREADS) lb_step_kernel(double* __restrict__ x, const double* __restrict__ x0,
                                                             const double* __restrict__ d, double t, size_t n) {
    if (VEC) {
        const size_t stride = 2 * (size_t)gridDim.x * LB_THREADS;
        size_t i = 2 * ((size_t)blockIdx.x * LB_THREADS + threadIdx.x);
        for (; i + 1 < n; i += stride) {
            const double2 a = ld2(x0 + i), dd = ld2(d + i);
            *reinterpret_cast<double2*>(x + i) = make_double2(fma(t, dd.x, a.x), fma(t, dd.y, a.y));
        }
        if (i < n) x[i] = fma(t, d[i], x0[i]);
    } else {
        const size_t stride = (size_t)gridDim.x * LB_THREADS;
        for (size_t i = (size_t)blockIdx.x * LB_THREADS + threadIdx.x; i < n; i += stride) x[i] = fma(t, d[i], x0[i]);
    }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int lb_blocks(size_t n, int per_sm) {
    size_t want = (n + 4 * LB_THREADS - 1) / (4 * LB_THREADS);          // at least two 16-byte loads per thread
    size_t cap = (size_t)svgpfa_sm_count() * per_sm;
    if (cap > (size_t)LB_MAXB) cap = LB_MAXB;
    if (want > cap) want = cap;
    return want < 1 ? 1 : (int)want;
}

}  // namespace

extern "C" uint64_t svgpfa_lbfgs_ws_doubles(void) { return (uint64_t)LB_MAXB * (LB_MAXV + LB_GROUP) * 3; }

extern "C" int svgpfa_lbfgs_multidot(const double* const* vecs_host, int32_t nv, const double* const* probes_host, int32_t np,
                                     uint64_t n, double* ws, double* out, void* stream) {
    if (!vecs_host || !probes_host || !ws || !out || nv < 1 || nv > LB_MAXV || np < 1 || np > 3)
        return svgpfa_set_error(SVGPFA_E_ARG, "lbfgs_multidot", cudaSuccess);
    LbPtrs vp, pp;
    const int ngroups = (nv + LB_GROUP - 1) / LB_GROUP;
    for (int v = 0; v < LB_MAXV; ++v) {
        vp.v[v] = v < nv ? vecs_host[v] : vecs_host[0];                  // padding of the last group: results never read
        if (!aligned16(vp.v[v])) return svgpfa_set_error(SVGPFA_E_ARG, "lbfgs_multidot: vectors must be 16-byte aligned", cudaSuccess);
    }
    for (int j = 0; j < LB_MAXV; ++j) {
        pp.v[j] = j < np ? probes_host[j] : probes_host[0];
        if (!aligned16(pp.v[j])) return svgpfa_set_error(SVGPFA_E_ARG, "lbfgs_multidot: vectors must be 16-byte aligned", cudaSuccess);
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = lb_blocks(n, 2);
    const dim3 grid(nb, ngroups);
    if (np == 1) lb_multidot_kernel<1><<<grid, LB_THREADS, 0, st>>>(vp, pp, n, ws);
    else if (np == 2) lb_multidot_kernel<2><<<grid, LB_THREADS, 0, st>>>(vp, pp, n, ws);
    else lb_multidot_kernel<3><<<grid, LB_THREADS, 0, st>>>(vp, pp, n, ws);
    const int nout = nv * np;
    lb_reduce_kernel<<<(nout + LB_RED_WARPS - 1) / LB_RED_WARPS, 32 * LB_RED_WARPS, 0, st>>>(ws, nb, nout, 0ull, out);
    SVGPFA_CHECK_LAUNCH("lbfgs_multidot");
    return SVGPFA_OK;
}

extern "C" int svgpfa_lbfgs_combine(double* d, const double* const* vecs_host, const double* coef_host, int32_t nv,
                                    int32_t accumulate, const double* g, uint64_t n, double* ws, double* out2, void* stream) {
    if (!d || !vecs_host || !coef_host || !g || !ws || !out2 || nv < 1 || nv > LB_MAXV || !aligned16(d) || !aligned16(g))
        return svgpfa_set_error(SVGPFA_E_ARG, "lbfgs_combine", cudaSuccess);
    LbPtrs vp;
    LbCoef cf;
    for (int v = 0; v < LB_MAXV; ++v) {
        vp.v[v] = v < nv ? vecs_host[v] : vecs_host[0];
        cf.c[v] = v < nv ? coef_host[v] : 0.0;
        if (!aligned16(vp.v[v])) return svgpfa_set_error(SVGPFA_E_ARG, "lbfgs_combine: vectors must be 16-byte aligned", cudaSuccess);
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = lb_blocks(n, 4);
    if (accumulate) lb_combine_kernel<true><<<nb, LB_THREADS, 0, st>>>(vp, cf, nv, d, g, n, ws);
    else lb_combine_kernel<false><<<nb, LB_THREADS, 0, st>>>(vp, cf, nv, d, g, n, ws);
    lb_reduce_kernel<<<1, 32 * LB_RED_WARPS, 0, st>>>(ws, nb, 2, 2ull, out2);
    SVGPFA_CHECK_LAUNCH("lbfgs_combine");
    return SVGPFA_OK;
}

extern "C" int svgpfa_lbfgs_stats(const double* a, const double* b, uint64_t n, double* ws, double* out4, void* stream) {
    if (!a || !ws || !out4 || !aligned16(a) || (b && !aligned16(b)))
        return svgpfa_set_error(SVGPFA_E_ARG, "lbfgs_stats", cudaSuccess);
    cudaStream_t st = (cudaStream_t)stream;
    const int nb = lb_blocks(n, 4);
    lb_stats_kernel<<<nb, LB_THREADS, 0, st>>>(a, b, n, ws);
    lb_reduce_kernel<<<1, 32 * LB_RED_WARPS, 0, st>>>(ws, nb, 4, 2ull | 8ull, out4);
    SVGPFA_CHECK_LAUNCH("lbfgs_stats");
    return SVGPFA_OK;
}

extern "C" int svgpfa_lbfgs_update(double* s, double* y, const double* d, double t, const double* g, double* g_prev,
                                   uint64_t n, void* stream) {
    if (!s || !y || !d || !g || !g_prev || !aligned16(s) || !aligned16(y) || !aligned16(d) || !aligned16(g) || !aligned16(g_prev))
        return svgpfa_set_error(SVGPFA_E_ARG, "lbfgs_update", cudaSuccess);
    if (n == 0) return SVGPFA_OK;
    lb_update_kernel<<<lb_blocks(n, 4), LB_THREADS, 0, (cudaStream_t)stream>>>(s, y, d, t, g, g_prev, n);
    SVGPFA_CHECK_LAUNCH("lbfgs_update");
    return SVGPFA_OK;
}

extern "C" int svgpfa_lbfgs_step(double* x, const double* x0, const double* d, double t, uint64_t n, void* stream) {
    if (!x || !x0 || !d) return svgpfa_set_error(SVGPFA_E_ARG, "lbfgs_step", cudaSuccess);
    if (n == 0) return SVGPFA_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (aligned16(x) && aligned16(x0) && aligned16(d)) lb_step_kernel<true><<<lb_blocks(n, 4), LB_THREADS, 0, st>>>(x, x0, d, t, n);
    else lb_step_kernel<false><<<lb_blocks(n, 4), LB_THREADS, 0, st>>>(x, x0, d, t, n);
    SVGPFA_CHECK_LAUNCH("lbfgs_step");
    return SVGPFA_OK;
}

/*
 * svgpfa_b200_probes.h -- measurement probes and test hooks, built into a SEPARATE library
 * (svgpfa_b200/libsvgpfa_b200_probes.so).  Nothing here is on the product path: the product library
 * (include/svgpfa_b200.h) carries no probe kernel.  Users: bench.py's roofline denominators (SURVEY.md §8d asks
 * the builder to MEASURE pi_fma / pi_exp / pi_sin on the box), tools/probe_*.py, tests/test_gpu_kernels.py.
 */
#ifndef SVGPFA_B200_PROBES_H
#define SVGPFA_B200_PROBES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

const char* svgpfa_probes_last_error(void);

/* FP64 pipe micro-benchmarks.  Launches `blocks` x 256 threads (kinds 20-23: 128 threads), each doing
 * `iters` x 8 operations; the caller times the launch with events.  out: `blocks*256` doubles (sink).
 *   kind 0: dependent-chain-free DFMA with constant operands      1: libdevice exp     2: libdevice sincospi
 *        3: the library's svgpfa_exp_neg                           4: mma.m8n8k4.f64 (8 per thread-iteration)
 *   5-9   : DFMA / DMMA mixes (are the two paths independent?)
 *   10-16 : issue-model probes (DFMA with three register operands, DADD, DMUL, interleaved integer / LDS)
 *   20-23 : the spike kernel's evaluation sequence in isolation (full / no moments / no table LDS / no spike LDS) */
int svgpfa_peak_probe(int32_t kind, int32_t blocks, int64_t iters, double* out, void* stream);

/* y_fast[i] = the library's exp for non-positive arguments, y_ref[i] = libdevice exp(x[i]). */
int svgpfa_exp_neg_eval(const double* x, double* y_fast, double* y_ref, int64_t n, void* stream);

/* y[i] = 2^(-min(w2[i], 2.61e5) / 256), the pre-scaled exponential of the spike kernels (w2 >= 0).
 * variant 0: degree-4 polynomial (<= 4.5e-16; what the kernels use); 3: degree-3 economised polynomial + I2F range
 * reduction (<= 2e-14; measured 7 % faster in tools/probe_eval.py, not adopted: see spike.cu). */
int svgpfa_exp2m_eval(const double* w2, double* y, int64_t n, int32_t variant, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SVGPFA_B200_PROBES_H */

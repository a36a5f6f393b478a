/*
 * svgpfa_b200.h -- C ABI of the B200-native svGPFA lower-bound hot path.
 *
 * The reference (joacorapela/svGPFA) is pure Python/PyTorch and has NO FFI; its seam for this
 * path is the duck-typed model protocol that SVEM_PyTorch drives.  This header
 * is the boundary a maintainer would bind from Python (ctypes/cffi) underneath that protocol:
 * plain pointers and sizes, no torch types.  Each entry point cites the reference code whose
 * arithmetic it replaces (paths relative to /root/reference/src/svGPFA/).
 *
 * Conventions
 *   - every pointer in svgpfa_buffers is a DEVICE pointer unless the field name ends in _host;
 *   - the library never allocates, frees or keeps caller memory across calls; it is re-entrant
 *     per stream; `stream` is a cudaStream_t passed as void*;
 *   - return value: 0 = launched OK, <0 = SVGPFA_E_* (bad argument / CUDA launch error).
 *     Numerical failure (non-positive-definite Kzz, the reference's torch.linalg.LinAlgError at
 *     utils/miscUtils.py:215) is reported asynchronously through buffers.info[0..3];
 *   - all arithmetic is IEEE float64.
 *
 * Data layout (R trials of this shard, N neurons, K latents, Q quadrature points, M_k inducing
 * points of latent k, P_k = M_k(M_k+1)/2):
 *   parameters & their gradients, "K-major", exactly the memory of the reference's list-of-K
 *   tensors laid end to end:   Z, m : [k][r][j]   offset R*moff[k] + r*M_k + j
 *                              cholvec: [k][r][p] offset R*poff[k] + r*P_k + p  (row-major tril order,
 *                                                 utils/miscUtils.py:135-139)
 *                              theta  : [thoff[k] + i]    C: [n][k]    d: [n]
 *   per-(trial, latent) workspace, "trial-major":
 *                              vectors  [r][moff[k] + j]            (length R*KM)
 *                              matrices [r][mmoff[k] + i*M_k + j]   (length R*MM)
 *   quadrature statistics      [r][q][k]  (the reference's (R,Q,K) layout, svPosteriorOnLatents.py:190-193)
 *   spikes                     spike_t[s] float64, trial-major then neuron-major -- the order of
 *                              PointProcessELL.__stackSpikeTimes (expectedLogLikelihood.py:157-173);
 *                              seg_off[r*N + n] .. seg_off[r*N + n + 1] = spikes of neuron n in trial r.
 */
#ifndef SVGPFA_B200_H
#define SVGPFA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVGPFA_ABI_VERSION 8
#define SVGPFA_MAX_M 64            /* inducing points per latent (north_star: M up to 64) */

enum { SVGPFA_KERNEL_EXPQUAD = 0, SVGPFA_KERNEL_PERIODIC = 1 };

/* which gradient groups a call must produce (honour needs_input_grad) */
enum {
    SVGPFA_GRAD_POSTERIOR = 1,     /* m, cholvec           (E-step,            svEM.py:218-223) */
    SVGPFA_GRAD_EMBEDDING = 2,     /* C, d                 (embedding M-step,  svEM.py:225-232) */
    SVGPFA_GRAD_KERNEL    = 4,     /* theta                (kernels M-step,    svEM.py:234-254) */
    SVGPFA_GRAD_INDLOCS   = 8,     /* Z                    (ind.-points M-step, svEM.py:256-264) */
    SVGPFA_GRAD_ALL       = 15,
    SVGPFA_REUSE_KZZ      = 16,    /* L, Li, logdet already valid for the current (Z, theta)   */
    SVGPFA_REUSE_SPIKE    = 32,    /* abar_spk already valid for the current (Z, theta, C)     */
    SVGPFA_REBUILD_PANELS = 64,    /* spike_method = PANEL: rebuild pm_tau for the trial range first (new spikes) */
    SVGPFA_REUSE_VQ       = 128    /* buffers.v_q already valid for the current (Z, theta): every E-step closure after the first */
};

/* How the spike-time term is evaluated (svgpfa_dims.spike_method).
 *   DIRECT  every (spike, latent, inducing point) kernel value is computed: S*K*M evaluations (svgpfa_spike_fwd_bwd).
 *   PANEL   the term only ever needs sums  sum_s c_s f(t_s)  of functions f in the span of kappa_k(. - z_j) and their
 *           parameter derivatives -- all analytic in t.  The time axis [pm_lo, pm_lo + pm_B*pm_w] is cut into pm_B
 *           panels; on each, f is replaced by its interpolant at SVGPFA_PM_P first-kind Chebyshev nodes, so
 *               sum_{s in (r,n)} f(t_s) = sum_i tau[r][n][i] f(t_i),   tau[r][n][i] = sum_{s in (r,n)} l_i(t_s)
 *           with l_i the Lagrange cardinal functions of the panel nodes: the spikes enter only through the STATIC
 *           "panel moments" tau (svgpfa_panel_moments, once per data set), and an evaluation costs
 *           R*NB*(K*M kernel values + two skinny GEMMs over N) with NB = pm_B*SVGPFA_PM_P nodes per trial instead of
 *           S_r spikes (config #5: 128 nodes against ~10^4 spikes per trial).  Interpolation error for the
 *           exponential-quadratic kernel with panel half-width h <= 0.625 lengthscale: <= 3e-14 relative (tests);
 *           the caller picks pm_B from the current hyper-parameters and falls back to DIRECT when it cannot. */
enum { SVGPFA_SPIKE_DIRECT = 1, SVGPFA_SPIKE_PANEL = 2 };
#define SVGPFA_PM_P 16             /* Chebyshev nodes per panel */

enum {
    SVGPFA_OK = 0, SVGPFA_E_ARG = -1, SVGPFA_E_CUDA = -2, SVGPFA_E_UNSUPPORTED = -3,
    SVGPFA_INFO_NOT_PD = 1        /* value of info[0] when a Kzz was not positive definite */
};

/* One row per latent, int32[8], device and host copies (see svgpfa_dims.desc_host). */
typedef struct svgpfa_latent_desc {
    int32_t ktype;   /* SVGPFA_KERNEL_*            */
    int32_t M;       /* inducing points            */
    int32_t moff;    /* sum of M over previous k   */
    int32_t mmoff;   /* sum of M*M over previous k */
    int32_t poff;    /* sum of P over previous k   */
    int32_t thoff;   /* sum of #params over previous k */
    int32_t P;       /* M(M+1)/2                   */
    int32_t nth;     /* 1 (expquad) or 2 (periodic)*/
} svgpfa_latent_desc;

typedef struct svgpfa_dims {
    int32_t R, N, K, Q;
    int32_t KM, MM, PP, TH;       /* sums over latents of M, M*M, P, #params */
    int32_t Mmax, n_ntiles;       /* n_ntiles = ceil(N / SVGPFA_EMBED_TN): neuron tiles of quad_embed */
    int64_t S;                    /* spikes in this shard */
    double  reg;                  /* prior-covariance regulariser (kernelsMatricesStore.py:113-116) */
    const svgpfa_latent_desc* desc_host;   /* HOST pointer, K rows */
    int32_t r0, rn;               /* trial range [r0, r0 + rn) the per-trial stages work on; rn = 0: all R trials.
                                     Sizes, strides and layouts are always those of the full shard (R); the host-buffer
                                     entry uses the range to pipeline copies and kernels over blocks of trials. */
    int32_t spike_chunks;         /* tuning: neuron ranges per trial in the spike kernel's grid; 0 = automatic
                                     (tests force 1: whole-trial ranges, several spike tiles per CTA) */
    int32_t quad_warps;           /* tuning: warps per CTA of the M <= 32 quadrature kernels (one 32-point tile per warp
                                     and pass); 0 = automatic */
    int32_t spike_method;         /* SVGPFA_SPIKE_*; 0 = DIRECT */
    int32_t pm_B;                 /* PANEL: panels per trial: 4, 8, 12, 16, 24 or 32 */
    double  pm_lo, pm_w;          /* PANEL: start of the first panel, panel width (same for every trial of the shard) */
} svgpfa_dims;

#define SVGPFA_NTRIALS(d) ((d)->rn ? (d)->rn : (d)->R)

#define SVGPFA_EMBED_TN 128

typedef struct svgpfa_buffers {
    /* ---- inputs ------------------------------------------------------------------------- */
    const svgpfa_latent_desc* desc;  /* K rows                                                  */
    const double* kscale;        /* [K][4]: scale^2, 1/lengthscaleScale, 1/periodScale, 0 (kernels.py:29-31,37,75-76) */
    const double* theta;         /* TH                                                          */
    const double* Z;             /* R*KM   K-major                                              */
    const double* m;             /* R*KM   K-major                                              */
    const double* cholvec;       /* R*PP   K-major                                              */
    const double* C;             /* N*K                                                         */
    const double* d;             /* N                                                           */
    const double* tq;            /* R*Q    quadrature nodes                                     */
    const double* wq;            /* R*Q    quadrature weights                                   */
    const double* spike_t;       /* S                                                           */
    const int64_t* seg_off;      /* R*N+1                                                       */
    const double* spike_cnt;     /* N      spikes of neuron n summed over the shard's trials    */
    /* ---- per-(trial,latent) workspace --------------------------------------------------- */
    double* L;                   /* R*MM  chol(Kzz)                                             */
    double* Li;                  /* R*MM  L^-1 (lower)                                          */
    double* X;                   /* R*MM  L^-1 Ls (lower)                                       */
    double* c;                   /* R*KM  L^-1 m                                                */
    double* alpha;               /* R*KM  Kzz^-1 m                                              */
    double* logdetL;             /* R*K   sum_i log L_ii                                        */
    double* kl_rk;               /* R*K                                                         */
    double* A_q;                 /* R*MM  sum_q varbar_q v_q v_q^T (lower incl. diag)           */
    double* abar_q;              /* R*KM  sum_q mubar_q k_q                                     */
    double* abar_spk;            /* R*KM  sum_s C[n_s,k] kappa(t_s, z_j)                        */
    double* dz_acc;              /* R*KM  dELBO/dZ contributions of the quadrature and spike terms */
    double* dth_part;            /* R*TH  per-trial dELBO/dtheta partials                       */
    /* ---- quadrature statistics ---------------------------------------------------------- */
    double* mu_q;                /* R*Q*K */
    double* var_q;               /* R*Q*K */
    double* v_q;                 /* R*KM*Q  [r][moff_k + j][q]  OPTIONAL (may be NULL): V = L^-1 kappa(Z, t_q) of every quadrature
                                    point, written by svgpfa_quad_latent_fwd and read back by svgpfa_quad_latent_bwd instead of
                                    being rebuilt (M <= 32, Q even; 20.5 GB at config #5).  Must be NULL when the buffers of a
                                    call do not belong to this Q (post-fit read-outs at other times) */
    double* mubar_part;          /* n_ntiles*R*K*Q  [tile][r][k][q] per-neuron-tile partials of dELBO/dmu_q */
    double* varbar_part;         /* n_ntiles*R*K*Q  same layout                                  */
    double* term1_part;          /* SVGPFA_TERM1_SLOTS partial sums of the intensity integral  */
    double* fin_part;            /* 3*SVGPFA_FIN_SLOTS per-block partials of svgpfa_finalize (KL, term1, term2), summed in
                                    fixed order: the bound is run-to-run reproducible given its inputs */
    /* ---- panel path of the spike-time term (spike_method = SVGPFA_SPIKE_PANEL) ----------- */
    double* pm_tau;              /* R*N*NB   [r][n][i]  static panel moments of every (trial, neuron) spike train      */
    double* pm_mun;              /* R*K*NB   [r][k][i]  latent means at the panel nodes                                */
    double* pm_mt;               /* R*K*NB   [r][k][i]  node weights  sum_n C[n,k] tau[r][n][i]                        */
    /* ---- cached-statistics path (embedding M-step) -------------------------------------- */
    double* mu_s;                /* S*K   latent means at spike times, [s][k]                   */
    double* gsum;                /* N*K   sum over the shard's spikes of neuron n of mu_s[s][k]: with cached statistics the
                                    spike part of the ELL is sum_{n,k} gsum[n][k] C[n][k], linear in C                */
    /* ---- outputs ------------------------------------------------------------------------ */
    double* shared;              /* SVGPFA_SHARED_HDR + N*K + N + TH: [elbo, ell, kl, term1, term2, status, r, k | dC | dd | dtheta]
                                    -- the buffer a multi-GPU caller all-reduces; status/r/k = info[0..2]
                                    as doubles, so that the sum over shards is > 0 iff some shard's Kzz failed and one
                                    small device->host copy of the header carries the bound AND the error state */
    double* gZ;                  /* R*KM  K-major */
    double* gm;                  /* R*KM  K-major */
    double* gcholvec;            /* R*PP  K-major */
    int32_t* info;               /* 4: [status, r, k, 0] */
} svgpfa_buffers;

#define SVGPFA_SHARED_HDR 8
#define SVGPFA_TERM1_SLOTS 4096
#define SVGPFA_FIN_SLOTS 1024

int svgpfa_abi_version(void);
const char* svgpfa_last_error(void);

/* (i)+(ii)  Kzz = kappa(Z,Z) + reg I, L = chol(Kzz), Li = L^-1, logdetL.
 * Replaces IndPointsLocsKMS.buildKernelsMatrices + IndPointsLocsKMS_Chol._invertKzz3D
 * (stats/kernelsMatricesStore.py:107-128) and miscUtils.chol3D (utils/miscUtils.py:209-216). */
int svgpfa_kzz_chol_fwd(const svgpfa_dims* dims, const svgpfa_buffers* buf, void* stream);

/* (ii)+(v)  Ls = unpack(cholvec), X = Li Ls, c = Li m, alpha = Li^T c, KL_rk.
 * Replaces SVPosteriorOnIndPointsChol.buildCov (stats/svPosteriorOnIndPoints.py:47-49,
 * utils/miscUtils.py:135-155), IndPointsLocsKMS_Chol.solveForLatent (kernelsMatricesStore.py:132-138)
 * and KLDivergence._evalSumAcrossTrials (stats/klDivergence.py:31-44). */
int svgpfa_indpoints_fwd(const svgpfa_dims* dims, const svgpfa_buffers* buf, void* stream);

/* (i)+(iii) latent posterior mean/variance at the Legendre quadrature points.
 * Replaces IndPointsLocsAndAllTimesKMS.buildKernelsMatrices (kernelsMatricesStore.py:186-195) and
 * SVPosteriorOnLatentsAllTimes.__computeMeansAndVarsGivenKernelMatrices (svPosteriorOnLatents.py:185-216);
 * Ktz (R,Q,M) is never materialised. */
int svgpfa_quad_latent_fwd(const svgpfa_dims* dims, const svgpfa_buffers* buf, void* stream);
/* the same from buffers.v_q (V = L^-1 kappa(Z, t_q), written by an earlier svgpfa_quad_latent_fwd for the same Z, theta):
 * mu_q = v_q . c, var_q = s2 - |v_q|^2 + |X^T v_q|^2 -- no kernel evaluation (svgpfa_elbo_grad with SVGPFA_REUSE_VQ);
 * falls back to svgpfa_quad_latent_fwd when the cache does not apply */
int svgpfa_quad_latent_fwd_cached(const svgpfa_dims* dims, const svgpfa_buffers* buf, void* stream);

/* (iii) embedding + exp-link intensity integral + its adjoints (dC, dd into `shared`, mubar/varbar partials).
 * Replaces LinearSVEmbeddingAllTimes._computeMeansAndVarsGivenSVPosteriorOnLatentsStats (svEmbedding.py:80-84),
 * PointProcessELLExpLink._getELinkValues (expectedLogLikelihood.py:205-208) and the weighted sum at
 * expectedLogLikelihood.py:122-131; eLinkValues (R,Q,N) is never materialised. */
int svgpfa_quad_embed_fwd_bwd(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, void* stream);

/* (iii) adjoint of svgpfa_quad_latent_fwd: A_q, abar_q and (if KERNEL|INDLOCS) dz_acc, dth_part (ADDED to: the
 * caller zeroes dz_acc and dth_part once per evaluation, svgpfa_elbo_grad does).
 * The reference obtains these from torch.autograd (svEM.py:281). */
int svgpfa_quad_latent_bwd(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, void* stream);

/* (i)+(iv) spike-time log-intensity term, value and every adjoint in one pass over the ragged CSR
 * spikes: abar_spk, dC (into `shared`), dz_acc, dth_part.
 * Replaces IndPointsLocsAndAssocTimesKMS.buildKernelsMatrices (kernelsMatricesStore.py:208-221),
 * SVPosteriorOnLatentsAssocTimes.__compute... (svPosteriorOnLatents.py:265-300; the unused variance
 * is not computed), LinearSVEmbeddingAssocTimes._compute... (svEmbedding.py:137-144) and
 * PointProcessELLExpLink._getELogLinkValues (expectedLogLikelihood.py:210-213). */
int svgpfa_spike_fwd_bwd(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, void* stream);

/* The same term and the same outputs through the panel moments (see SVGPFA_SPIKE_PANEL above).
 *   svgpfa_panel_moments       : pm_tau from spike_t / seg_off for the trial range (static: once per data set and
 *                                panelisation)
 *   svgpfa_spike_panel_fwd_bwd : pm_mun (latent means at the nodes), pm_mt = C^T tau, dC += tau mun^T, then abar_spk,
 *                                dz_acc, dth_part from the node weights -- kernel values at R*K*M*NB points
 *   svgpfa_panel_neuron_sums   : gsum[n][k] = sum_{spikes of neuron n} mu_k(t_s) for the cached-statistics path
 *                                (embedding M-step) without materialising the S x K spike-time means */
int svgpfa_panel_moments(const svgpfa_dims* dims, const svgpfa_buffers* buf, void* stream);
int svgpfa_panel_neuron_sums(const svgpfa_dims* dims, const svgpfa_buffers* buf, void* stream);
int svgpfa_spike_panel_fwd_bwd(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, void* stream);

/* (ii)+(v) adjoints through alpha, c, X, the KL term and the Cholesky factorisation: gm, gcholvec,
 * gZ, dth_part.  The reference obtains these from torch.autograd. */
int svgpfa_indpoints_bwd(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, void* stream);

/* reductions into `shared`: elbo = -term1 + term2 - KL, dtheta = sum_r dth_part, dd += spike counts.
 * Replaces SVLowerBound.eval (svLowerBound.py:47-54) and the sums at expectedLogLikelihood.py:132-134,
 * klDivergence.py:18-29. */
int svgpfa_finalize(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, void* stream);

/* The whole path: the eight calls above in order (zeroing accumulators first). One unit of work of
 * the benchmark = this call with flags = SVGPFA_GRAD_ALL. */
int svgpfa_elbo_grad(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, void* stream);

/* (iv) cached-statistics path of the embedding M-step (svEM.py:225-232):
 *   svgpfa_spike_latent_means   : mu_s[s][k] = kappa_k(t_s, Z_kr) . alpha_kr
 *                                 (SVPosteriorOnLatentsAssocTimes.computeMeansAndVars, mean part)
 *   svgpfa_cached_ell_fwd_bwd   : ELL(C, d | mu_q, var_q, mu_s) and dC, dd into `shared`
 *                                 (PointProcessELL.evalSumAcrossTrialsAndNeurons with
 *                                  svPosteriorOnLatentsStats, expectedLogLikelihood.py:107-135).
 *                                 The spike part is the HBM-bound ragged gather over mu_s into gsum; it does not depend
 *                                 on (C, d), so flags = SVGPFA_REUSE_SPIKE skips it when gsum already holds the sums of
 *                                 these statistics (every closure evaluation of an embedding M-step after the first). */
int svgpfa_spike_latent_means(const svgpfa_dims* dims, const svgpfa_buffers* buf, void* stream);
int svgpfa_cached_ell_fwd_bwd(const svgpfa_dims* dims, const svgpfa_buffers* buf, uint32_t flags, void* stream);

/* Post-fit read-outs at arbitrary times.  The latent statistics at the times come from
 * svgpfa_quad_latent_fwd with dims.Q = number of times per trial and buffers.tq / mu_q / var_q pointing at the times and
 * the outputs (SVPosteriorOnLatentsAllTimes.predict, svPosteriorOnLatents.py:57-77); this call turns them into
 *   e_mean = mu C^T + d, e_var = var (C^T)^2   (LinearSVEmbeddingAllTimes.predict, svEmbedding.py:86-92)
 *   cif    = exp(e_mean + e_var / 2)            (PointProcessELLExpLink.computeExpectedPosteriorCIFs,
 *                                                expectedLogLikelihood.py:62-73)
 * each (R, Q, N) row-major; any of the three outputs may be NULL. */
int svgpfa_embed_predict(const svgpfa_dims* dims, const svgpfa_buffers* buf, double* e_mean, double* e_var, double* cif,
                         void* stream);

/* Host helper: (r,n) segment offsets from per-segment spike counts, and the per-spike neuron index the
 * reference builds (expectedLogLikelihood.py:168-172), for the bit-exact indexing check.
 * counts_host[R*N] -> seg_off_host[R*N+1]; neuron_index_host may be NULL, else S entries (int64). */
int svgpfa_build_segments_host(int32_t R, int32_t N, const int64_t* counts_host,
                               int64_t* seg_off_host, int64_t* neuron_index_host);

/* End-to-end entry with HOST buffers (pinned for overlap; pageable works): copies parameters and spikes to the
 * device, runs the stages of svgpfa_elbo_grad and copies `shared`, gZ, gm, gcholvec, info back; all of it is
 * ordered on `stream` from the caller's point of view (the call returns without synchronising).  `dev` supplies the
 * device-side buffers (same struct, device pointers); every *_host array mirrors the device one. */
typedef struct svgpfa_host_io {
    const double* theta_host; const double* Z_host; const double* m_host; const double* cholvec_host;
    const double* C_host; const double* d_host; const double* tq_host; const double* wq_host;
    const double* spike_t_host; const int64_t* seg_off_host; const double* spike_cnt_host;
    double* shared_host; double* gZ_host; double* gm_host; double* gcholvec_host; int32_t* info_host;
    int32_t copy_static;      /* 1: also copy tq, wq, spikes, segments (first call); 0: parameters only */
    int32_t n_blocks;         /* blocks of trials the copies and kernels are pipelined over (copy-in and copy-out
                                 streams owned by the library run under the kernels of the neighbouring blocks);
                                 0 = automatic (R / 312, at most 16), 1 = everything in order on `stream` */
} svgpfa_host_io;
int svgpfa_elbo_grad_host(const svgpfa_dims* dims, const svgpfa_buffers* dev, const svgpfa_host_io* io,
                          uint32_t flags, void* stream);

/* Destroys the side streams and events svgpfa_elbo_grad_host created for the CALLING host thread (one set per device,
 * created on first use); the thread must have drained its streams.  Optional: a process that exits need not call it. */
int svgpfa_release_thread_resources(void);

/* Measurement hook (bench.py only).
 * svgpfa_set_stage_events: when `events` is non-NULL the next svgpfa_elbo_grad calls on this host thread
 *   record events[0] before the first stage and events[i+1] after stage i (stages in the order of
 *   SVGPFA_STAGE_*; a skipped stage records its event immediately), so that a caller can read every
 *   kernel's device time inside its own timed region.  `events` = SVGPFA_N_STAGES+1 cudaEvent_t handles.
 *   Pass NULL to switch recording off.
 * (The FP64 peak probes and the exp test hooks live in a separate library: include/svgpfa_b200_probes.h.) */
enum { SVGPFA_STAGE_KZZ_CHOL = 0, SVGPFA_STAGE_INDPOINTS_FWD, SVGPFA_STAGE_QUAD_LATENT_FWD, SVGPFA_STAGE_QUAD_EMBED,
       SVGPFA_STAGE_QUAD_LATENT_BWD, SVGPFA_STAGE_SPIKE, SVGPFA_STAGE_INDPOINTS_BWD, SVGPFA_STAGE_FINALIZE,
       SVGPFA_N_STAGES };
int svgpfa_set_stage_events(void** events);

/* ---- Vector primitives of the device-resident, shard-aware L-BFGS (svgpfa_b200/lbfgs.py).
 * The reference optimises every ECM step with torch.optim.LBFGS (stats/svEM.py:218-294: one instance per step, closure
 * = -eval + backward).  The replacement keeps the optimiser state on the device and runs the two-loop recursion in
 * coefficient space: an iteration touches the stored (s, y) pairs twice -- svgpfa_lbfgs_multidot (new rows of their
 * Gram matrix) and svgpfa_lbfgs_combine (the search direction) -- and under trial sharding exchanges one small
 * all-reduce of those rows.  All vectors are device pointers to n float64, 16-byte aligned unless noted; `ws` is a
 * device workspace of svgpfa_lbfgs_ws_doubles() float64 (per-block partials, combined in block order: results are
 * run-to-run reproducible); outputs are device pointers (the caller copies them to the host when it needs to branch).
 *   multidot : out[i * np + j] = vecs[i] . probes[j],   1 <= nv <= SVGPFA_LBFGS_MAX_VECS, 1 <= np <= 3;
 *              vecs_host / probes_host are HOST arrays of device pointers
 *   combine  : d = (accumulate ? d : 0) + sum_i coef_host[i] vecs[i];  out2 = [g . d, max|d|]  (histories longer than
 *              SVGPFA_LBFGS_MAX_VECS are combined in several calls)
 *   stats    : out4 = [a . b, max|a|, sum|a|, max|b|]  (b may be NULL)
 *   update   : s = t d,  y = g - g_prev,  g_prev = g     (torch/optim/lbfgs.py: "do lbfgs update (update memory)")
 *   step     : x = x0 + t d   (x, x0, d need only 8-byte alignment: x is a slice of a packed parameter buffer) */
#define SVGPFA_LBFGS_MAX_VECS 64
#define SVGPFA_LBFGS_MAX_BLOCKS 592
uint64_t svgpfa_lbfgs_ws_doubles(void);
int svgpfa_lbfgs_multidot(const double* const* vecs_host, int32_t nv, const double* const* probes_host, int32_t np,
                          uint64_t n, double* ws, double* out, void* stream);
int svgpfa_lbfgs_combine(double* d, const double* const* vecs_host, const double* coef_host, int32_t nv,
                         int32_t accumulate, const double* g, uint64_t n, double* ws, double* out2, void* stream);
int svgpfa_lbfgs_stats(const double* a, const double* b, uint64_t n, double* ws, double* out4, void* stream);
int svgpfa_lbfgs_update(double* s, double* y, const double* d, double t, const double* g, double* g_prev,
                        uint64_t n, void* stream);
int svgpfa_lbfgs_step(double* x, const double* x0, const double* d, double t, uint64_t n, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SVGPFA_B200_H */

"""``B200SVLowerBound``: the model object ``SVEM_PyTorch.maximize`` drives, with the method
names, argument meaning and error behaviour of the reference's ``SVLowerBound``
(``/root/reference/src/svGPFA/stats/svLowerBound.py:6-120``; protocol table in SURVEY.md §8b),
whose arithmetic runs in the hand-written CUDA library behind ``include/svgpfa_b200.h``.

Differences a user can observe, all deliberate:
  * parameters live on the GPU; the getters return the leaf tensors the optimiser mutates in
    place (views into one packed buffer per parameter group, so no gather/scatter per step);
  * ``buildKernelsMatrices()`` only invalidates caches; matrices are rebuilt lazily inside
    ``eval()`` when (Z, theta) actually changed (tensor version counters);
  * value and gradients are produced by ONE fused forward+backward pass (the bound is a sum,
    its upstream gradient is a scalar applied in ``backward``);
  * a non-positive-definite Kzz raises ``torch.linalg.LinAlgError`` from ``eval()``
    (the reference raises it from ``buildKernelsMatrices``, utils/miscUtils.py:215).  The error state
    travels with the bound in the 8-double header of the result buffer, which is copied to pinned host
    memory asynchronously: a forward-only ``eval()`` checks it before returning; an evaluation that will
    be differentiated (an optimiser closure) is checked when its copy has landed -- at the latest on the
    next call into the model -- so the host never stalls between ``eval()`` and ``backward()``.
    ``checkErrors()`` forces the check;
  * with a ``process_group`` (one rank per GPU, trials sharded) an evaluation all-reduces the packed
    ``[elbo.. | dC | dd | dtheta]`` buffer, except while ONLY per-trial leaves (m, cholVecs, Z) require
    gradients: the bound is separable over trials given (C, d, theta), so each rank then optimises its own
    block with no collective inside the closure (``shard_mode``; SURVEY.md §8e option (i)).
There is no CPU path: without a CUDA device or without the built library every entry raises.
"""
from __future__ import annotations

import ctypes
import math
import warnings

import numpy as np
import torch

from . import _cabi, ingest, sharding
from .kernels import kernel_spec

_F64 = torch.float64


def _tril_size(M):
    return M * (M + 1) // 2


class _LowerBoundFn(torch.autograd.Function):
    """ELBO (or ELL from cached statistics) as one differentiable node over every leaf."""

    @staticmethod
    def forward(ctx, model, cached_stats, *leaves):
        K = model._K
        need = ctx.needs_input_grad[2:]
        flags = 0
        if any(need[0:2 * K]):
            flags |= _cabi.GRAD_POSTERIOR
        if any(need[2 * K:2 * K + 2]):
            flags |= _cabi.GRAD_EMBEDDING
        if any(need[2 * K + 2:3 * K + 2]):
            flags |= _cabi.GRAD_KERNEL
        if any(need[3 * K + 2:4 * K + 2]):
            flags |= _cabi.GRAD_INDLOCS
        if cached_stats is not None:
            flags &= _cabi.GRAD_EMBEDDING
            shared = model._run_cached(cached_stats)
            gZ = gm = gcv = None
        else:
            shared, gZ, gm, gcv = model._run(flags)
        ctx.model = model
        ctx.flags = flags
        ctx.bufs = (shared, gZ, gm, gcv)
        model._last_shared = shared
        ctx.d_shape = leaves[2 * K + 1].shape
        return shared[0].clone()

    @staticmethod
    def backward(ctx, grad_out):
        model, flags = ctx.model, ctx.flags
        model._poll_errors(block=False)
        shared, gZ, gm, gcv = ctx.bufs
        K, R, N = model._K, model._R, model._N
        need = ctx.needs_input_grad[2:]
        grads = [None] * (4 * K + 2)
        if flags & _cabi.GRAD_POSTERIOR:
            sm, sc = gm * grad_out, gcv * grad_out
            for k in range(K):
                M, P = model._M[k], model._P[k]
                if need[k]:
                    grads[k] = sm[R * model._moff[k]:R * (model._moff[k] + M)].view(R, M, 1)
                if need[K + k]:
                    grads[K + k] = sc[R * model._poff[k]:R * (model._poff[k] + P)].view(R, P, 1)
        if flags & (_cabi.GRAD_EMBEDDING | _cabi.GRAD_KERNEL):
            ss = shared * grad_out
            h = _cabi.SHARED_HDR
            if need[2 * K]:
                grads[2 * K] = ss[h:h + N * K].view(N, K)
            if need[2 * K + 1]:
                grads[2 * K + 1] = ss[h + N * K:h + N * K + N].view(ctx.d_shape)
            if flags & _cabi.GRAD_KERNEL:
                th = ss[h + N * K + N:]
                for k in range(K):
                    if need[2 * K + 2 + k]:
                        grads[2 * K + 2 + k] = th[model._thoff[k]:model._thoff[k] + model._nth[k]]
        if flags & _cabi.GRAD_INDLOCS:
            sz = gZ * grad_out
            for k in range(K):
                if need[3 * K + 2 + k]:
                    M = model._M[k]
                    grads[3 * K + 2 + k] = sz[R * model._moff[k]:R * (model._moff[k] + M)].view(R, M, 1)
        return (None, None, *grads)


class _LazySpikeMeans:
    """Per-trial (S_r, K) latent means at the spike times, computed on first access (sequence protocol of the list
    the reference returns, expectedLogLikelihood.py:141-147).  Valid while the model's parameters are the ones the
    statistics were computed for."""

    def __init__(self, model, offsets, snapshot):
        self._model, self._off, self._snapshot, self._rows = model, offsets, snapshot, None

    def _materialise(self):
        if self._rows is None:
            if self._model._param_snapshot() != self._snapshot:
                raise RuntimeError("the spike-time means of these statistics were not materialised before the model's "
                                   "parameters changed; read stats['assocTimes'][0] right after computing the statistics")
            mu_s = self._model._spike_time_means()
            self._rows = [mu_s[self._off[r]:self._off[r + 1]] for r in range(len(self._off) - 1)]
        return self._rows

    def __len__(self):
        return len(self._off) - 1

    def __getitem__(self, i):
        return self._materialise()[i]

    def __iter__(self):
        return iter(self._materialise())


class _LazySpikeVars(_LazySpikeMeans):
    """Per-trial (S_r, K) latent variances at the spike times, computed on first access.  The exponential link never
    reads them (expectedLogLikelihood.py:210-213), so the lower-bound path does not compute them; a caller that does
    (PointProcessELLQuad, expectedLogLikelihood.py:215-255, or a diagnostic) gets the values of
    SVPosteriorOnLatentsAssocTimes.computeMeansAndVars (svPosteriorOnLatents.py:265-300) from the quadrature kernels
    run at the spike times."""

    def _materialise(self):
        if self._rows is None:
            if self._model._param_snapshot(with_cov=True) != self._snapshot:
                raise RuntimeError("the spike-time variances of these statistics were not materialised before the "
                                   "model's parameters changed")
            self._rows = self._model._spike_time_vars(self._off)
        return self._rows


class B200SVLowerBound:
    _PENDING_SLOTS = 8
    _stats_serial = 0

    def __init__(self, kernels=None, device=None, process_group=None, check_errors=True, shard_mode="auto",
                 ind_points_cov_rep="chol", kzz_inv_method="chol"):
        if ind_points_cov_rep not in ("chol", "rank1_plus_diag"):
            raise ValueError("ind_points_cov_rep must be 'chol' or 'rank1_plus_diag'")
        if kzz_inv_method not in ("chol", "pinv"):
            raise ValueError("kzz_inv_method must be 'chol' or 'pinv'")
        # how Kzz^-1 is applied (stats/svGPFAModelFactory.py:25-27).  "pinv" is the reference's IndPointsLocsKMS_PInv
        # (kernelsMatricesStore.py:146-159: torch.linalg.pinv(Kzz, rcond=1e-15) @ x).  For a Kzz of full numerical rank --
        # which kappa(Z, Z) + reg I is whenever its Cholesky factorisation exists -- the pseudo-inverse IS the inverse, and
        # so is its derivative: the variant runs through the same kernels (the reference's two variants agree to 3e-16 on
        # the bound and 3e-14 on the gradients on the committed fixture).  What is NOT reproduced is the truncation of a
        # numerically singular Kzz: there the Cholesky fails and the error below is raised.
        self._kzz_inv_method = kzz_inv_method
        # how the variational covariance S_kr is parameterised (stats/svGPFAModelFactory.py:29-32):
        #   "chol"             Cholesky vectors, S = Ls Ls^T          (SVPosteriorOnIndPointsChol, the default)
        #   "rank1_plus_diag"  S = q q^T + diag(d^2)                  (SVPosteriorOnIndPointsRank1PlusDiag)
        # The kernels consume Cholesky vectors; with (q, d) they are DERIVED on the device at every evaluation
        # (batched torch.linalg.cholesky of q q^T + diag(d^2): plumbing, O(R K M^3) against the path's O(R K Q M^2))
        # and the gradient the kernels return for them is carried on to (q, d) by autograd.
        self._cov_rep = ind_points_cov_rep
        self._device = torch.device(device) if device is not None else None
        self._pg = process_group
        if shard_mode not in ("auto", "reduce", "local"):
            raise ValueError("shard_mode must be 'auto', 'reduce' or 'local'")
        self.shard_mode = shard_mode
        self.v_cache = True               # False: the quadrature adjoint rebuilds V instead of reading it back (saves R*KM*Q*8 bytes)
        self._check_errors = check_errors
        self._pending = []                 # slots of the header copies in flight, oldest first
        self._pinned = None
        self._leaf_list = None
        self._next_slot = 0
        self._spike_chunks = 0             # tuning / tests: neuron ranges per trial in the spike kernel (0 = automatic)
        self._quad_warps = 0               # tuning: warps per CTA of the quadrature kernels (0 = automatic)
        self.spike_method = "auto"         # "auto" | "direct" | "panel": how the spike-time term is evaluated
        self._pm = None                    # state of the panel path: dict(B, lo, w, theta_version, built)
        self._gsum_key = None              # which cached spike-time means the per-neuron sums in `gsum` belong to
        self._kernels = None
        self._reg = None
        self._params_set = False
        self._spikes_set = False
        self._quad_set = False
        self._ready = False
        self._kzz_key = self._vq_key = None
        self._spike_key = None
        self._bufs = None
        if kernels is not None:
            self.setKernels(kernels)

    # ------------------------------------------------------------------ device / library
    def _dev(self):
        if self._device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("svgpfa_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
            self._device = torch.device("cuda", torch.cuda.current_device())
        if self._device.type != "cuda":
            raise RuntimeError("svgpfa_b200 only runs on CUDA devices; there is no CPU fallback")
        return self._device

    def _to_dev(self, x):
        if isinstance(x, torch.Tensor):
            return x.detach().to(device=self._dev(), dtype=_F64)
        return torch.as_tensor(np.asarray(x), dtype=_F64).to(self._dev())

    # ------------------------------------------------------------------ setters (svLowerBound.py:13-45,83-99)
    def setKernels(self, kernels):
        self._kernels = list(kernels)
        self._ready = False
        self._kzz_key = self._spike_key = self._vq_key = None
        if self._params_set:               # kernels replaced after the parameters: re-derive what depends on them
            if len(self._kernels) != self._K:
                raise ValueError("inconsistent number of latents in kernels / initial_params")
            self._bind_kernels()

    def _bind_kernels(self):
        """Kernel-dependent metadata (type, scale^2, parameter scalings) and the aliasing of every kernel object's
        parameter tensor with its slice of the packed buffer."""
        K, dev = self._K, self._dev()
        specs = [kernel_spec(kern) for kern in self._kernels]
        nth = [2 if s[0] == _cabi.KERNEL_PERIODIC else 1 for s in specs]
        if getattr(self, "_nth", nth) != nth:
            raise ValueError("kernel types do not match the packed kernel parameters; call setInitialParams again")
        self._nth = nth
        for k, kern in enumerate(self._kernels):
            if hasattr(kern, "setParams"):
                kern.setParams(self._theta[k])
        desc = (_cabi.LatentDesc * K)()
        for k in range(K):
            desc[k] = _cabi.LatentDesc(specs[k][0], self._M[k], self._moff[k], self._mmoff[k], self._poff[k],
                                       self._thoff[k], self._P[k], self._nth[k])
        self._desc_host = desc
        self._desc_dev = torch.frombuffer(bytearray(bytes(desc)), dtype=torch.int32).to(dev)
        self._kscale_host = [[s[1], s[2], s[3], 0.0] for s in specs]
        self._kscale = torch.tensor(self._kscale_host, dtype=_F64).to(dev).contiguous()

    def _make_views(self):
        """The leaf tensors the getters hand out: views of the packed K-major buffers (the optimiser's in-place
        updates need no gather, and the kernels read the packed buffers directly)."""
        R, K = self._R, self._K
        self._Z = [self._Zbuf[R * self._moff[k]:R * self._moff[k + 1]].view(R, self._M[k], 1) for k in range(K)]
        self._m = [self._mbuf[R * self._moff[k]:R * self._moff[k + 1]].view(R, self._M[k], 1) for k in range(K)]
        self._cv = [self._cvbuf[R * self._poff[k]:R * self._poff[k + 1]].view(R, self._P[k], 1) for k in range(K)]
        self._theta = [self._thbuf[self._thoff[k]:self._thoff[k + 1]] for k in range(K)]
        if getattr(self, "_cov_rep", "chol") == "rank1_plus_diag":
            self._qv = [self._qvbuf[R * self._moff[k]:R * self._moff[k + 1]].view(R, self._M[k], 1) for k in range(K)]
            self._qd = [self._qdbuf[R * self._moff[k]:R * self._moff[k + 1]].view(R, self._M[k], 1) for k in range(K)]
        self._leaf_list = None

    def _rank1(self):
        return getattr(self, "_cov_rep", "chol") == "rank1_plus_diag"

    def _derive_cholvecs(self):
        """Rank-1-plus-diagonal parameterisation: Cholesky vectors of S = q q^T + diag(d^2) (row-major tril order,
        miscUtils.py:135-139) as tensors ATTACHED to the autograd graph of (q, d); their values are also written
        into the packed buffer the kernels read (svPosteriorOnIndPoints.py:103-115, buildCov)."""
        out = []
        for k in range(self._K):
            q, dg = self._qv[k][:, :, 0], self._qd[k][:, :, 0]
            S = q.unsqueeze(2) * q.unsqueeze(1) + torch.diag_embed(dg * dg)
            L = torch.linalg.cholesky(S)
            ti = torch.tril_indices(self._M[k], self._M[k], device=L.device)
            cv = L[:, ti[0], ti[1]].unsqueeze(-1)
            with torch.no_grad():
                self._cv[k].copy_(cv)
            out.append(cv)
        return out

    def _refresh_leaf_list(self):
        self._leaf_list = list(self._m) + list(self._cv) + [self._C, self._d] + list(self._theta) + list(self._Z)

    def setInitialParams(self, initial_params):
        """``initial_params`` as produced by ``svGPFA.utils.initUtils.getParamsAndKernelsTypes``
        (utils/initUtils.py:468-481).  The tensors are copied into packed device buffers; the
        getters return leaf views of those buffers."""
        if self._kernels is None:
            raise RuntimeError("setKernels must be called before setInitialParams")
        pol = initial_params["posterior_on_latents"]
        mean = pol["posterior_on_ind_points"]["mean"]
        if self._rank1():
            # (R, M, 1) each (svPosteriorOnIndPoints.py:91-94); the Cholesky vectors are derived, see _derive_cholvecs
            qsv, qsd = pol["posterior_on_ind_points"]["qSVec0"], pol["posterior_on_ind_points"]["qSDiag0"]
            chol = [torch.zeros(int(q.shape[0]), _tril_size(int(q.shape[1])), 1, dtype=_F64) for q in qsv]
        else:
            chol = pol["posterior_on_ind_points"]["cholVecs"]
        kms = pol["kernels_matrices_store"]
        theta0, Z0 = kms["kernels_params0"], kms["inducing_points_locs0"]
        C0, d0 = initial_params["embedding"]["C0"], initial_params["embedding"]["d0"]
        K = len(mean)
        if not (len(chol) == K and len(theta0) == K and len(Z0) == K and len(self._kernels) == K):
            raise ValueError("inconsistent number of latents in initial_params / kernels")
        R = int(Z0[0].shape[0])
        self._K, self._R = K, R
        self._M = [int(Z0[k].shape[1]) for k in range(K)]
        if max(self._M) > _cabi.MAX_M:
            raise ValueError(f"at most {_cabi.MAX_M} inducing points per latent are supported")
        self._P = [_tril_size(M) for M in self._M]
        for k in range(K):
            if tuple(mean[k].shape) != (R, self._M[k], 1) or tuple(chol[k].shape) != (R, self._P[k], 1) \
                    or tuple(Z0[k].shape) != (R, self._M[k], 1):
                raise ValueError(f"latent {k}: expected mean (R,M,1), cholVecs (R,P,1), locs (R,M,1)")
        specs = [kernel_spec(kern) for kern in self._kernels]
        self._nth = [2 if s[0] == _cabi.KERNEL_PERIODIC else 1 for s in specs]
        for k in range(K):
            if int(theta0[k].numel()) != self._nth[k]:
                raise ValueError(f"latent {k}: kernel expects {self._nth[k]} parameters")
        cum = lambda xs: [int(v) for v in np.concatenate([[0], np.cumsum(xs)])]
        self._moff, self._poff, self._thoff = cum(self._M), cum(self._P), cum(self._nth)
        self._mmoff = cum([M * M for M in self._M])
        self._KM, self._PP, self._TH, self._MM = self._moff[-1], self._poff[-1], self._thoff[-1], self._mmoff[-1]
        pack = lambda xs: torch.cat([self._to_dev(x).reshape(-1) for x in xs]).contiguous()
        self._Zbuf, self._mbuf, self._cvbuf, self._thbuf = pack(Z0), pack(mean), pack(chol), pack(theta0)
        if self._rank1():
            for k in range(K):
                if tuple(qsv[k].shape) != (R, self._M[k], 1) or tuple(qsd[k].shape) != (R, self._M[k], 1):
                    raise ValueError(f"latent {k}: expected qSVec0 and qSDiag0 of shape (R,M,1)")
            self._qvbuf, self._qdbuf = pack(qsv), pack(qsd)
        self._make_views()
        self._C = self._to_dev(C0).contiguous().clone()
        self._d = self._to_dev(d0).contiguous().clone()
        self._N = int(self._C.shape[0])
        if self._C.shape[1] != K or self._d.numel() != self._N:
            raise ValueError("C must be (N,K) and d must have N entries")
        self._bind_kernels()
        self._refresh_leaf_list()
        self._params_set = True
        self._ready = False
        self._kzz_key = self._spike_key = self._vq_key = None
        if self._rank1():
            with torch.no_grad():
                self._derive_cholvecs()

    def setMeasurements(self, measurements):
        """``measurements[r][n]`` = spike times of neuron n in trial r (list / array / tensor, float32
        or float64).  Stacked trial-major, neuron-major, within-neuron order kept -- the order of
        ``PointProcessELL.__stackSpikeTimes`` (stats/expectedLogLikelihood.py:157-173)."""
        times, counts = ingest.stack_spike_times(measurements)
        self.setMeasurementsFlat(times, counts)

    def setMeasurementsFlat(self, spike_times, spike_counts):
        """Same data already stacked: ``spike_times`` (S,) in the order described above and
        ``spike_counts`` (R, N).  Accepts host or device arrays."""
        dev = self._dev()
        if isinstance(spike_counts, torch.Tensor):
            counts_dev = spike_counts.to(device=dev, dtype=torch.int64).contiguous()
            R, N = counts_dev.shape
            seg = torch.zeros(R * N + 1, dtype=torch.int64, device=dev)
            torch.cumsum(counts_dev.reshape(-1), 0, out=seg[1:])
            cnt = counts_dev.sum(0).to(_F64)
        else:
            counts = np.ascontiguousarray(np.asarray(spike_counts, dtype=np.int64))
            R, N = counts.shape
            seg_host = np.empty(R * N + 1, dtype=np.int64)
            _cabi.check(_cabi.lib().svgpfa_build_segments_host(
                R, N, counts.ctypes.data, seg_host.ctypes.data, None), "build_segments_host")
            seg = torch.from_numpy(seg_host).to(dev)
            cnt = torch.from_numpy(counts.sum(0).astype(np.float64)).to(dev)
        if isinstance(spike_times, torch.Tensor):
            st = spike_times.detach().to(device=dev, dtype=_F64).contiguous()
        else:
            st = torch.from_numpy(np.ascontiguousarray(np.asarray(spike_times)).astype(np.float64)).to(dev)
        self._S = int(st.numel())
        if self._S:
            lo_hi = torch.stack([st.min(), st.max()]).cpu()
            self._t_lo, self._t_hi = float(lo_hi[0]), float(lo_hi[1])
        else:
            self._t_lo, self._t_hi = 0.0, 1.0
        self._pm = None
        self._seg_off, self._spike_t, self._spike_cnt = seg.contiguous(), st.reshape(-1), cnt.contiguous()
        self._spike_R, self._spike_N = int(R), int(N)
        self._spikes_set = True
        self._ready = False
        self._spike_key = None

    def setELLCalculationParams(self, eLLCalculationParams):
        self._tq = self._to_dev(eLLCalculationParams["leg_quad_points"]).reshape(
            eLLCalculationParams["leg_quad_points"].shape[0], -1).contiguous()
        self._wq = self._to_dev(eLLCalculationParams["leg_quad_weights"]).reshape(self._tq.shape).contiguous()
        self._Q = int(self._tq.shape[1])
        self._quad_set = True
        self._ready = False

    def setPriorCovRegParam(self, priorCovRegParam):
        self._reg = float(priorCovRegParam)
        self._kzz_key = self._spike_key = self._vq_key = None
        if self._ready:
            self._dims.reg = self._reg         # the value the kernels read

    def setParamsAndData(self, measurements, initial_params, eLLCalculationParams, priorCovRegParam):
        self.setMeasurements(measurements=measurements)
        self.setInitialParams(initial_params=initial_params)
        self.setELLCalculationParams(eLLCalculationParams=eLLCalculationParams)
        self.setPriorCovRegParam(priorCovRegParam=priorCovRegParam)
        self.buildKernelsMatrices()

    # ------------------------------------------------------------------ getters (svLowerBound.py:101-114)
    def getSVPosteriorOnIndPointsParams(self):
        if self._rank1():                  # mean, qSVec, qSDiag (svPosteriorOnIndPoints.py:96-101)
            return list(self._m) + list(self._qv) + list(self._qd)
        return list(self._m) + list(self._cv)

    def getSVEmbeddingParams(self):
        return [self._C, self._d]

    def getKernels(self):
        return self._kernels

    def getKernelsParams(self):
        return list(self._theta)

    def getIndPointsLocs(self):
        return list(self._Z)

    # ------------------------------------------------------------------ buffers
    def _prepare(self):
        if self._rank1() and self._params_set:
            # the packed Cholesky vectors follow (q, d) for the entries that do not go through _apply (latent statistics,
            # read-outs)
            key = (self._qvbuf._version, self._qdbuf._version)
            if getattr(self, "_cov_key", None) != key:
                with torch.no_grad():
                    self._derive_cholvecs()
                self._cov_key = key
        if self._ready:
            return
        if not (self._params_set and self._spikes_set and self._quad_set and self._reg is not None):
            raise RuntimeError("model is not fully specified: call setKernels, setInitialParams, "
                               "setMeasurements, setELLCalculationParams and setPriorCovRegParam first")
        R, N, K, Q = self._R, self._N, self._K, self._Q
        if self._spike_R != R or self._spike_N != N:
            raise ValueError(f"measurements are ({self._spike_R} trials, {self._spike_N} neurons) "
                             f"but parameters are ({R}, {N})")
        if self._tq.shape[0] != R:
            raise ValueError("leg_quad_points must have one row per trial")
        _cabi.lib()
        dev = self._dev()
        e = lambda *shape: torch.empty(*shape, dtype=_F64, device=dev)
        n_ntiles = (N + _cabi.EMBED_TN - 1) // _cabi.EMBED_TN
        self._n_ntiles = n_ntiles
        ws = dict(
            L=e(R * self._MM), Li=e(R * self._MM), X=e(R * self._MM), c=e(R * self._KM), alpha=e(R * self._KM),
            logdetL=e(R * K), kl_rk=e(R * K), A_q=torch.zeros(R * self._MM, dtype=_F64, device=dev),
            abar_q=torch.zeros(R * self._KM, dtype=_F64, device=dev),
            abar_spk=torch.zeros(R * self._KM, dtype=_F64, device=dev),
            dz_acc=torch.zeros(R * self._KM, dtype=_F64, device=dev),
            dth_part=torch.zeros(R * self._TH, dtype=_F64, device=dev),
            mu_q=e(R * Q * K), var_q=e(R * Q * K), mubar_part=e(n_ntiles * R * Q * K),
            varbar_part=e(n_ntiles * R * Q * K),
            term1_part=torch.zeros(_cabi.TERM1_SLOTS, dtype=_F64, device=dev),
            fin_part=torch.zeros(3 * _cabi.FIN_SLOTS, dtype=_F64, device=dev),
            gsum=torch.zeros(max(N * K, 1), dtype=_F64, device=dev),
            info=torch.zeros(4, dtype=torch.int32, device=dev))
        # V = L^-1 kappa(Z, t_q) of every quadrature point, handed from the forward to the adjoint quadrature kernel
        # (include/svgpfa_b200.h: buffers.v_q) when it fits comfortably: 20.5 GB at config #5
        vq_bytes = 8 * R * self._KM * Q
        if getattr(self, "v_cache", True) and max(self._M) <= 64 and Q % 2 == 0 and vq_bytes > 0:
            free, _ = torch.cuda.mem_get_info(dev)
            if vq_bytes <= 0.4 * free:
                ws["v_q"] = e(R * self._KM * Q)
        self._ws = ws
        self._shared_len = _cabi.SHARED_HDR + N * K + N + self._TH
        dims = _cabi.Dims(R=R, N=N, K=K, Q=Q, KM=self._KM, MM=self._MM, PP=self._PP, TH=self._TH,
                          Mmax=max(self._M), n_ntiles=n_ntiles, S=self._S, reg=self._reg,
                          desc_host=ctypes.cast(self._desc_host, ctypes.POINTER(_cabi.LatentDesc)),
                          spike_chunks=int(self._spike_chunks), quad_warps=int(self._quad_warps))
        self._dims = dims
        b = _cabi.Buffers()
        ptr = lambda t: ctypes.c_void_p(t.data_ptr())
        b.desc, b.kscale, b.theta = ptr(self._desc_dev), ptr(self._kscale), ptr(self._thbuf)
        b.Z, b.m, b.cholvec = ptr(self._Zbuf), ptr(self._mbuf), ptr(self._cvbuf)
        b.C, b.d, b.tq, b.wq = ptr(self._C), ptr(self._d), ptr(self._tq), ptr(self._wq)
        b.spike_t, b.seg_off, b.spike_cnt = ptr(self._spike_t), ptr(self._seg_off), ptr(self._spike_cnt)
        for name, t in ws.items():
            setattr(b, name, ptr(t))
        b.mu_s = None
        b.gsum = ptr(ws["gsum"])
        b.pm_tau = b.pm_mun = b.pm_mt = None
        self._bufs = b
        self._pm = None
        self._gsum_key = None
        self._ready = True
        self._kzz_key = self._spike_key = self._vq_key = None

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self._dev()).cuda_stream)

    def _param_versions(self):
        return (self._Zbuf._version, self._thbuf._version, self._reg)

    # ------------------------------------------------------------------ the hot path
    def buildKernelsMatrices(self):
        """Reference semantics (svLowerBound.py:77-78): everything derived from (theta, Z) is
        recomputed from their current values at the next evaluation."""
        self._kzz_key = self._vq_key = None
        self._spike_key = None

    # ------------------------------------------------------------------ spike-term method (include/svgpfa_b200.h)
    PANEL_COUNTS = (4, 8, 12, 16, 24, 32)
    PANEL_BETA = {_cabi.KERNEL_EXPQUAD: 0.625, _cabi.KERNEL_PERIODIC: 0.55}

    def _panels_required(self, theta_host):
        """Smallest panel count for which the 16-node interpolation of every latent's kernel (and of its parameter
        derivatives) is accurate to ~3e-14 of the sums it replaces: panel half-width <= beta x (scale of variation).
        Exponential-quadratic: the length scale, beta = 0.625.  Periodic, exp(-2 sin^2(pi d/p)/l^2): p min(l, 1)/(2 pi)
        -- a Gaussian of that length scale near d = 0 for small l, harmonics of the period for large l -- beta = 0.55.
        Calibrated in tools/panel_accuracy.py (numpy, long sums of random weights), checked on the GPU in
        tests/test_gpu_panel.py."""
        span = max(self._t_hi - self._t_lo, 1e-300)
        need = 1
        for k in range(self._K):
            ktype = self._desc_host[k].ktype
            th = theta_host[self._thoff[k]:self._thoff[k + 1]]
            ell = abs(float(th[0])) * float(self._kscale_host[k][1])
            if ktype == _cabi.KERNEL_PERIODIC:
                ell = min(ell, 1.0) * abs(float(th[1])) * float(self._kscale_host[k][2]) / (2.0 * math.pi)
            need = max(need, math.ceil(span / (2.0 * self.PANEL_BETA[ktype] * max(ell, 1e-300)) - 1e-9))
        return need

    def _select_spike_method(self, theta_host=None, build=True):
        """Chooses DIRECT or PANEL for the current hyper-parameters (they fix the panel count) and (re)builds the
        static panel moments when the panelisation changed.  Costs one small device->host copy when theta changed.
        ``theta_host``: take the hyper-parameters from this host array instead (host-buffer entry: the device copy
        is about to be overwritten); ``build=False``: the caller rebuilds the moments itself (new spikes)."""
        dims, b, dev = self._dims, self._bufs, self._dev()
        ver = self._thbuf._version if theta_host is None else ("host", tuple(np.asarray(theta_host).tolist()))
        if self._pm is not None and self._pm["theta_version"] == ver:
            return
        if self.spike_method == "direct" or self._S == 0 or self._N == 0:
            dims.spike_method = _cabi.SPIKE_DIRECT
            self._pm = dict(B=0, theta_version=ver)
            return
        need = self._panels_required(self._thbuf.detach().cpu().numpy() if theta_host is None else np.asarray(theta_host))
        B = next((c for c in self.PANEL_COUNTS if c >= need), None)
        use = B is not None
        if use and self.spike_method == "auto":
            NB, KM, Kp = B * _cabi.PM_P, self._KM, 8 * ((self._K + 7) // 8)
            panel_cost = self._R * NB * (30.0 * KM + 2.0 * self._N * Kp + 45.0 * self._N)
            # the direct kernel gives a lane to every (latent, inducing point) pair of a trial, 128 pairs per CTA: with
            # few pairs (config #2: 27) most lanes idle, and its cost is that of the padded pair count
            direct_cost = 13.0 * self._S * 128.0 * math.ceil(KM / 128.0)
            tau_bytes = 8.0 * self._R * self._N * NB
            use = (1.5 * panel_cost < direct_cost and self._K <= 40 and self._N <= 8000
                   and tau_bytes <= 0.3 * torch.cuda.get_device_properties(dev).total_memory)
        if not use:
            if self.spike_method == "panel":
                raise RuntimeError(f"spike_method='panel' needs {need} panels (> 32) for the current kernel parameters")
            dims.spike_method = _cabi.SPIKE_DIRECT
            self._pm = dict(B=0, theta_version=ver)
            return
        old = self._pm or {}
        if old.get("B") != B:
            NB = B * _cabi.PM_P
            lo = self._t_lo
            w = (self._t_hi - self._t_lo) / B * (1.0 + 1e-12) if self._t_hi > self._t_lo else 1.0 / B
            self._ws["pm_tau"] = torch.empty(self._R * self._N * NB, dtype=_F64, device=dev)
            self._ws["pm_mun"] = torch.empty(self._R * self._K * NB, dtype=_F64, device=dev)
            self._ws["pm_mt"] = torch.empty(self._R * self._K * NB, dtype=_F64, device=dev)
            for name in ("pm_tau", "pm_mun", "pm_mt"):
                setattr(b, name, self._ws[name].data_ptr())
            dims.spike_method, dims.pm_B, dims.pm_lo, dims.pm_w = _cabi.SPIKE_PANEL, B, lo, w
            if build:
                with torch.cuda.device(dev):
                    _cabi.check(_cabi.lib().svgpfa_panel_moments(ctypes.byref(dims), ctypes.byref(b), self._stream()),
                                "panel_moments")
            self._spike_key = None
        dims.spike_method = _cabi.SPIKE_PANEL
        self._pm = dict(B=B, theta_version=ver)

    def _run(self, flags):
        """One fused value+gradient pass (svgpfa_elbo_grad).  Returns (shared, gZ, gm, gcholvec)."""
        self._prepare()
        self._select_spike_method()
        dev = self._dev()
        R = self._R
        kz_key = self._param_versions()
        sp_key = kz_key + (self._C._version,)
        call_flags = flags
        if self._kzz_key == kz_key:
            call_flags |= _cabi.REUSE_KZZ
        has_vq = "v_q" in self._ws
        if has_vq and getattr(self, "_vq_key", None) == kz_key:
            call_flags |= _cabi.REUSE_VQ             # V of the quadrature points depends on (Z, theta) only: E-step closures
        if self._spike_key == sp_key and not (flags & (_cabi.GRAD_KERNEL | _cabi.GRAD_INDLOCS | _cabi.GRAD_EMBEDDING)):
            call_flags |= _cabi.REUSE_SPIKE
        b = self._bufs
        shared = torch.empty(self._shared_len, dtype=_F64, device=dev)
        b.shared = shared.data_ptr()
        gZ = gm = gcv = None
        if flags & _cabi.GRAD_INDLOCS:
            gZ = torch.empty(R * self._KM, dtype=_F64, device=dev)
            b.gZ = gZ.data_ptr()
        if flags & _cabi.GRAD_POSTERIOR:
            gm = torch.empty(R * self._KM, dtype=_F64, device=dev)
            gcv = torch.empty(R * self._PP, dtype=_F64, device=dev)
            b.gm, b.gcholvec = gm.data_ptr(), gcv.data_ptr()
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().svgpfa_elbo_grad(ctypes.byref(self._dims), ctypes.byref(b),
                                                     call_flags, self._stream()), "elbo_grad")
        self._kzz_key, self._spike_key = kz_key, sp_key
        self._vq_key = kz_key if has_vq else None
        self._finish(shared, flags)
        return shared, gZ, gm, gcv

    # ------------------------------------------------------------------ exchange step and error channel
    def _reduces(self, flags):
        """Does this evaluation all-reduce the packed buffer?  """
        if self._pg is None or self.shard_mode == "local":
            return False
        if self.shard_mode == "reduce":
            return True
        return sharding.evaluation_is_reduced(flags)

    def _finish(self, shared, flags):
        if self._reduces(flags):
            # the one exchange step of the path: [elbo.., status | dC | dd | dtheta] summed over trial shards; the
            # status travels in the header, so every rank sees a failed Cholesky of any shard and raises with it
            sharding.all_reduce_shared(shared, self._pg)
        if self._check_errors:
            self._enqueue_check(shared)

    def _enqueue_check(self, shared):
        """Asynchronous copy of the 8-double header [elbo, ell, kl, term1, term2, status, r, k] to pinned memory."""
        if self._pinned is None:
            self._pinned = torch.zeros(self._PENDING_SLOTS, _cabi.SHARED_HDR, dtype=_F64).pin_memory()
            self._events = [torch.cuda.Event() for _ in range(self._PENDING_SLOTS)]
        if len(self._pending) == self._PENDING_SLOTS:
            self._poll_errors(block=True, limit=1)
        slot = self._next_slot
        self._next_slot = (slot + 1) % self._PENDING_SLOTS
        self._pinned[slot].copy_(shared[:_cabi.SHARED_HDR], non_blocking=True)
        self._events[slot].record(torch.cuda.current_stream(self._dev()))
        self._pending.append(slot)

    def _poll_errors(self, block, limit=None):
        """Examines the header copies that have landed (all of them when ``block``)."""
        done = 0
        while self._pending and (limit is None or done < limit):
            slot = self._pending[0]
            ev = self._events[slot]
            if block:
                ev.synchronize()
            elif not ev.query():
                break
            self._pending.pop(0)
            done += 1
            hdr = self._pinned[slot]
            if float(hdr[_cabi.SHARED_STATUS]) > 0.0:
                self._pending.clear()
                self._kzz_key = self._spike_key = self._vq_key = None
                where = (f"Kzz of trial {int(hdr[6])}, latent {int(hdr[7])} is" if float(hdr[_cabi.SHARED_STATUS]) == 1.0
                         else "a Kzz of more than one trial shard is")
                hint = (" (kzz_inv_method='pinv': the truncated pseudo-inverse of a numerically singular Kzz is not "
                        "reproduced)" if getattr(self, "_kzz_inv_method", "chol") == "pinv" else "")
                raise torch.linalg.LinAlgError(f"linalg.cholesky: {where} not positive-definite{hint}")
            if math.isinf(float(hdr[0])):
                warnings.warn("infinity lower bound detected")       # svLowerBound.py:51-53

    def checkErrors(self):
        """Blocks until every evaluation issued so far has finished and raises what it reported."""
        self._poll_errors(block=True)

    def _leaves(self):
        return self._leaf_list

    def _apply(self, cached_stats):
        self._poll_errors(block=False)
        leaves = self._leaf_list
        if self._rank1():                  # the Cholesky-vector slots carry tensors derived from (q, d)
            K = self._K
            leaves = leaves[:K] + self._derive_cholvecs() + leaves[2 * K:]
        differentiated = torch.is_grad_enabled() and any(p.requires_grad for p in leaves)
        out = _LowerBoundFn.apply(self, cached_stats, *leaves)
        if not differentiated:
            # nobody will call backward(): the caller reads the value next, so report a failure here
            # (reference semantics: the exception leaves the call that hit it)
            self._poll_errors(block=True)
        return out

    def eval(self):
        """ELL - KL as a 0-dim float64 tensor, differentiable w.r.t. every leaf that currently has
        ``requires_grad=True`` (svLowerBound.py:47-54)."""
        self._prepare()
        return self._apply(None)

    # ------------------------------------------------------------------ host-buffer entry (bench.py "e2e")
    def makeHostIO(self, pin=True):
        """Host mirrors (pinned by default) of every input and output of ``svgpfa_elbo_grad_host``,
        initialised from the current device state."""
        self._prepare()
        h = lambda t: (t.detach().cpu().pin_memory() if pin else t.detach().cpu())
        z = lambda n, dt=_F64: (torch.zeros(n, dtype=dt).pin_memory() if pin else torch.zeros(n, dtype=dt))
        R = self._R
        return dict(theta=h(self._thbuf), Z=h(self._Zbuf), m=h(self._mbuf), cholvec=h(self._cvbuf),
                    C=h(self._C), d=h(self._d), tq=h(self._tq), wq=h(self._wq), spike_t=h(self._spike_t),
                    seg_off=h(self._seg_off), spike_cnt=h(self._spike_cnt),
                    shared=z(self._shared_len), gZ=z(R * self._KM), gm=z(R * self._KM), gcholvec=z(R * self._PP),
                    info=z(4, torch.int32))

    def evalAndGradHost(self, io, flags=_cabi.GRAD_ALL, copy_static=True, n_blocks=0):
        """(Cholesky-vector parameterisation only.)
        One unit of work through the C ABI with HOST buffers: host->device copies of the inputs,
        ``svgpfa_elbo_grad``, device->host copies of the bound and the gradients, all on the current
        stream; returns after the stream has drained.  Returns (elbo, h2d_bytes, d2h_bytes)."""
        if self._rank1():
            raise NotImplementedError("the host-buffer entry takes Cholesky vectors: ind_points_cov_rep='chol'")
        self._prepare()
        self._select_spike_method(theta_host=io["theta"].numpy(), build=not copy_static)
        dev = self._dev()
        R = self._R
        b = _cabi.Buffers.from_buffer_copy(self._bufs)
        shared = torch.empty(self._shared_len, dtype=_F64, device=dev)
        gZ = torch.empty(R * self._KM, dtype=_F64, device=dev)
        gm = torch.empty(R * self._KM, dtype=_F64, device=dev)
        gcv = torch.empty(R * self._PP, dtype=_F64, device=dev)
        b.shared, b.gZ, b.gm, b.gcholvec = shared.data_ptr(), gZ.data_ptr(), gm.data_ptr(), gcv.data_ptr()
        hio = _cabi.HostIO()
        for name in ("theta", "Z", "m", "cholvec", "C", "d", "tq", "wq", "spike_t", "seg_off", "spike_cnt",
                     "shared", "gZ", "gm", "gcholvec", "info"):
            setattr(hio, name + "_host", io[name].data_ptr())
        hio.copy_static = 1 if copy_static else 0
        hio.n_blocks = int(n_blocks)          # 0: automatic; 1: no copy/compute overlap
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().svgpfa_elbo_grad_host(ctypes.byref(self._dims), ctypes.byref(b), ctypes.byref(hio),
                                                          flags, self._stream()), "elbo_grad_host")
            if self._pg is not None and self.shard_mode != "local":
                # the one exchange step, on the device; the reduced buffer replaces the local one on the host
                sharding.all_reduce_shared(shared, self._pg)
                io["shared"].copy_(shared, non_blocking=True)
            torch.cuda.current_stream(dev).synchronize()
        self._kzz_key = self._spike_key = self._vq_key = None
        if self._pm is not None:
            self._pm["theta_version"] = None        # the device parameters were replaced: re-derive at the next eval()
        nb = lambda *names: sum(io[n].numel() * io[n].element_size() for n in names)
        h2d = nb("theta", "Z", "m", "cholvec", "C", "d")
        if copy_static:
            h2d += nb("tq", "wq", "spike_t", "seg_off", "spike_cnt")
        d2h = nb("shared", "info")
        if flags & _cabi.GRAD_INDLOCS:
            d2h += nb("gZ")
        if flags & _cabi.GRAD_POSTERIOR:
            d2h += nb("gm", "gcholvec")
        if int(io["info"][0]) == _cabi.INFO_NOT_PD or float(io["shared"][_cabi.SHARED_STATUS]) > 0.0:
            raise torch.linalg.LinAlgError("linalg.cholesky: Kzz is not positive-definite")
        return float(io["shared"][0]), h2d, d2h

    # ------------------------------------------------------------------ embedding M-step (svEM.py:225-232)
    def _spike_time_means(self):
        """(S, K) latent means at the spike times for the CURRENT parameters (direct kernel evaluations)."""
        dev = self._dev()
        b = _cabi.Buffers.from_buffer_copy(self._bufs)
        mu_s = torch.empty(max(self._S, 1) * self._K, dtype=_F64, device=dev)
        b.mu_s = mu_s.data_ptr()
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().svgpfa_spike_latent_means(ctypes.byref(self._dims), ctypes.byref(b), self._stream()))
        return mu_s[:self._S * self._K].view(self._S, self._K)

    def _spike_time_vars(self, off, budget_bytes=1 << 30):
        """[(S_r, K) variances at the spike times of trial r]: svgpfa_quad_latent_fwd over blocks of trials with the
        (padded) spike times of a trial in place of its quadrature nodes."""
        dev = self._dev()
        R, K = self._R, self._K
        counts = [off[r + 1] - off[r] for r in range(R)]
        rows = [None] * R
        spikes = self._spike_t.to(_F64) if self._spike_t.dtype != _F64 else self._spike_t
        lib = _cabi.lib()
        r0 = 0
        while r0 < R:
            T, r1 = max(counts[r0], 1), r0 + 1
            while r1 < R and max(T, counts[r1]) * (r1 + 1 - r0) * K * 16 <= budget_bytes:
                T = max(T, counts[r1])
                r1 += 1
            n = r1 - r0
            t = torch.zeros(n, T, dtype=_F64, device=dev)
            for i in range(n):
                if counts[r0 + i]:
                    t[i, :counts[r0 + i]] = spikes[off[r0 + i]:off[r0 + i + 1]]
            mu = torch.empty(n * T * K, dtype=_F64, device=dev)
            var = torch.empty(n * T * K, dtype=_F64, device=dev)
            dims = _cabi.Dims.from_buffer_copy(self._dims)
            dims.Q, dims.r0, dims.rn = T, r0, n
            b = _cabi.Buffers.from_buffer_copy(self._bufs)
            # the kernels index these arrays by the GLOBAL trial number: bias the pointers by the block's first trial
            b.tq = t.data_ptr() - 8 * r0 * T
            b.mu_q, b.var_q = mu.data_ptr() - 8 * r0 * T * K, var.data_ptr() - 8 * r0 * T * K
            b.v_q = None                                       # the V cache belongs to the quadrature grid
            with torch.cuda.device(dev):
                _cabi.check(lib.svgpfa_quad_latent_fwd(ctypes.byref(dims), ctypes.byref(b), self._stream()))
            var = var.view(n, T, K)
            for i in range(n):
                rows[r0 + i] = var[i, :counts[r0 + i]]
            r0 = r1
        return rows

    def computeSVPosteriorOnLatentsStats(self):
        """Latent posterior statistics at quadrature points (mean, var: (R,Q,K)) and at spike times (means; the
        variances, which the exponential link never reads, are a lazy sequence computed on first access).  Layout follows
        expectedLogLikelihood.py:141-147; the per-trial spike tensors are views of one (S,K) buffer.

        With cached statistics the spike part of the expected log-likelihood is linear in C with coefficients
        G[n,k] = sum over the spikes of neuron n of mu_k(t_s); when the panel path is active G comes straight from the
        panel moments and the (S,K) array (33 GB at config #5) is only built if somebody reads it
        (``stats["assocTimes"][0]`` is then a lazy sequence)."""
        self._prepare()
        self._select_spike_method()
        dev = self._dev()
        b, lib = self._bufs, _cabi.lib()
        kz_key = self._param_versions()
        panel = self._dims.spike_method == _cabi.SPIKE_PANEL
        with torch.cuda.device(dev):
            if self._kzz_key != kz_key:
                self._ws["info"].zero_()
                _cabi.check(lib.svgpfa_kzz_chol_fwd(ctypes.byref(self._dims), ctypes.byref(b), self._stream()))
                self._kzz_key = kz_key
            _cabi.check(lib.svgpfa_indpoints_fwd(ctypes.byref(self._dims), ctypes.byref(b), self._stream()))
            _cabi.check(lib.svgpfa_quad_latent_fwd(ctypes.byref(self._dims), ctypes.byref(b), self._stream()))
            self._vq_key = kz_key if "v_q" in self._ws else None
            if panel:
                _cabi.check(lib.svgpfa_panel_neuron_sums(ctypes.byref(self._dims), ctypes.byref(b), self._stream()))
        R, Q, K = self._R, self._Q, self._K
        mu_q = self._ws["mu_q"].view(R, Q, K).clone()
        var_q = self._ws["var_q"].view(R, Q, K).clone()
        if self._check_errors and int(self._ws["info"][0].item()) == _cabi.INFO_NOT_PD:
            raise torch.linalg.LinAlgError("linalg.cholesky: Kzz is not positive-definite")
        B200SVLowerBound._stats_serial += 1
        serial = B200SVLowerBound._stats_serial
        stats = {"allTimes": (mu_q, var_q), "_b200_id": serial}
        off = self._seg_off[::self._N].cpu().tolist() if self._N else [0] * (R + 1)
        if panel:
            stats["_b200_gsum"] = self._ws["gsum"].clone()
            self._gsum_key = (serial, None)
            stats["assocTimes"] = (_LazySpikeMeans(self, off, self._param_snapshot()),
                                   _LazySpikeVars(self, off, self._param_snapshot(with_cov=True)))
        else:
            mu_s = self._spike_time_means()
            stats["_b200_mu_s"] = mu_s
            stats["assocTimes"] = ([mu_s[off[r]:off[r + 1]] for r in range(R)],
                                   _LazySpikeVars(self, off, self._param_snapshot(with_cov=True)))
        return stats

    def _param_snapshot(self, with_cov=False):
        snap = (self._Zbuf._version, self._thbuf._version, self._mbuf._version, self._reg)
        return snap + (self._cvbuf._version,) if with_cov else snap

    def evalELLSumAcrossTrialsAndNeurons(self, svPosteriorOnLatentsStats=None):
        """Expected log-likelihood only (no KL).  With cached statistics it is a function of (C, d)
        alone (svLowerBound.py:72-75)."""
        self._prepare()
        if svPosteriorOnLatentsStats is None:
            svPosteriorOnLatentsStats = self.computeSVPosteriorOnLatentsStats()
        return self._apply(svPosteriorOnLatentsStats)

    def _run_cached(self, stats):
        dev = self._dev()
        mu_q, var_q = stats["allTimes"]
        mu_q = self._to_dev(mu_q).contiguous()
        var_q = self._to_dev(var_q).contiguous()
        b = _cabi.Buffers.from_buffer_copy(self._bufs)
        shared = torch.empty(self._shared_len, dtype=_F64, device=dev)
        b.shared, b.mu_q, b.var_q = shared.data_ptr(), mu_q.data_ptr(), var_q.data_ptr()
        # the per-neuron sums of the spike-time means do not depend on (C, d): gathered once per set of statistics
        serial = stats.get("_b200_id")            # a serial number, not an address: addresses are reused by the allocator
        gsum, mu_s = stats.get("_b200_gsum"), None
        if gsum is not None:                      # panel path: the sums came with the statistics
            key = (serial, None)
            if self._gsum_key != key:
                self._ws["gsum"].copy_(gsum)
            flags = _cabi.REUSE_SPIKE
        else:
            mu_s = stats.get("_b200_mu_s")
            if mu_s is None:
                mu_s = torch.cat([self._to_dev(x) for x in stats["assocTimes"][0]], 0)
            mu_s = self._to_dev(mu_s).contiguous()
            b.mu_s = mu_s.data_ptr()
            key = None if serial is None else (serial, mu_s._version)
            flags = _cabi.REUSE_SPIKE if (key is not None and self._gsum_key == key) else 0
        with torch.cuda.device(dev):
            _cabi.check(_cabi.lib().svgpfa_cached_ell_fwd_bwd(ctypes.byref(self._dims), ctypes.byref(b), flags,
                                                              self._stream()), "cached_ell_fwd_bwd")
        self._gsum_key = key
        self._finish(shared, _cabi.GRAD_EMBEDDING)
        self._cached_keepalive = (mu_q, var_q, mu_s)
        return shared

    # ------------------------------------------------------------------ post-fit read-outs
    def _predict(self, times, want_embedding=False, want_cif=False):
        self._prepare()
        dev = self._dev()
        t = self._to_dev(times).reshape(self._R, -1).contiguous()
        T = int(t.shape[1])
        e = lambda n: torch.empty(self._R * T * n, dtype=_F64, device=dev)
        mu, var = e(self._K), e(self._K)
        dims = _cabi.Dims.from_buffer_copy(self._dims)
        dims.Q = T
        b = _cabi.Buffers.from_buffer_copy(self._bufs)
        b.tq, b.mu_q, b.var_q = t.data_ptr(), mu.data_ptr(), var.data_ptr()
        b.v_q = None                                           # the V cache belongs to the quadrature grid
        lib = _cabi.lib()
        out = [mu.view(self._R, T, self._K), var.view(self._R, T, self._K)]
        with torch.cuda.device(dev):
            if self._kzz_key != self._param_versions():
                self._ws["info"].zero_()
                _cabi.check(lib.svgpfa_kzz_chol_fwd(ctypes.byref(self._dims), ctypes.byref(self._bufs), self._stream()))
                self._kzz_key = self._param_versions()
            _cabi.check(lib.svgpfa_indpoints_fwd(ctypes.byref(self._dims), ctypes.byref(self._bufs), self._stream()))
            _cabi.check(lib.svgpfa_quad_latent_fwd(ctypes.byref(dims), ctypes.byref(b), self._stream()))
            if want_embedding or want_cif:
                em = e(self._N) if want_embedding else None
                ev = e(self._N) if want_embedding else None
                cif = e(self._N) if want_cif else None
                ptr = lambda x: None if x is None else ctypes.c_void_p(x.data_ptr())
                _cabi.check(lib.svgpfa_embed_predict(ctypes.byref(dims), ctypes.byref(b), ptr(em), ptr(ev), ptr(cif),
                                                     self._stream()))
                shape = (self._R, T, self._N)
                out += [None if x is None else x.view(shape) for x in (em, ev, cif)]
        if self._check_errors and int(self._ws["info"][0].item()) == _cabi.INFO_NOT_PD:
            self._kzz_key = self._vq_key = None
            raise torch.linalg.LinAlgError("linalg.cholesky: Kzz is not positive-definite")
        return out

    def predictLatents(self, times):
        """Posterior mean and variance of the latents at ``times`` (R, T, 1), each (R, T, K)
        (svLowerBound.py:116-117 -> svPosteriorOnLatents.py:57-77).  Forward-only call of the quadrature kernels."""
        mu, var = self._predict(times)[:2]
        return mu, var

    def predictEmbedding(self, times):
        """Embedding mean / variance at ``times``, each (R, T, N) (svLowerBound.py:119-120 -> svEmbedding.py:86-92)."""
        out = self._predict(times, want_embedding=True)
        return out[2], out[3]

    def computeExpectedPosteriorCIFs(self, times):
        """exp(mean + var/2) of the embedding per trial and neuron: ``answer[r][n]`` is a (T,) tensor
        (svLowerBound.py:64-66 -> expectedLogLikelihood.py:62-73)."""
        cif = self._predict(times, want_cif=True)[4]
        return [[cif[r, :, n] for n in range(cif.shape[2])] for r in range(cif.shape[0])]

    # ------------------------------------------------------------------ pickling (svEM.py:89-92,175-181)
    # ``pickle`` does not preserve storage sharing between tensors (only torch.save does), so the leaf views -- and
    # the kernel objects' parameter tensors, which alias the same packed buffer -- are NOT part of the state: they
    # are rebuilt from the packed buffers on load, with their requires_grad flags.
    _TRANSIENT = ("_bufs", "_dims", "_desc_host", "_desc_dev", "_ws", "_cached_keepalive", "_pg", "_Z", "_m", "_cv",
                  "_qv", "_qd", "_cov_key", "_theta", "_leaf_list", "_pending", "_pinned", "_events", "_next_slot", "_kernels", "_last_shared", "_pm", "_gsum_key")

    def __getstate__(self):
        self._poll_errors(block=True)
        st = {k: v for k, v in self.__dict__.items() if k not in self._TRANSIENT}
        st["_ready"] = False
        st["_kzz_key"] = st["_spike_key"] = st["_vq_key"] = None
        if self._params_set:
            st["_leaf_requires_grad"] = [bool(p.requires_grad) for p in self._leaf_list]
            if self._rank1():
                st["_qd_requires_grad"] = [bool(p.requires_grad) for p in list(self._qv) + list(self._qd)]
        if self._kernels is not None:
            kernels = []
            for kern in self._kernels:                 # kernel objects without their (aliasing) parameter tensor
                kst = dict(kern.__dict__)
                kst.pop("_params", None)
                kernels.append((type(kern), kst))
            st["_kernel_states"] = kernels
        return st

    def __setstate__(self, st):
        flags = st.pop("_leaf_requires_grad", None)
        qd_flags = st.pop("_qd_requires_grad", None)
        kernels = st.pop("_kernel_states", None)
        self.__dict__.update(st)
        self._pg = None
        self._bufs = None
        self._pending, self._pinned, self._next_slot = [], None, 0
        self._kernels = None
        self._leaf_list = None
        self._last_shared = None
        self._pm = None
        self._gsum_key = None
        if kernels is not None:
            self._kernels = []
            for cls, kst in kernels:
                kern = cls.__new__(cls)
                kern.__dict__.update(kst)
                self._kernels.append(kern)
        if self._params_set:
            self._make_views()
            self._bind_kernels()
            self._refresh_leaf_list()
            for p, flag in zip(self._leaf_list, flags or []):
                p.requires_grad_(flag)
            if self._rank1():
                for p, flag in zip(list(self._qv) + list(self._qd), qd_flags or []):
                    p.requires_grad_(flag)


indPointsCovRank1PlusDiag, indPointsCovChol = 100000, 100001      # the reference's constants (svGPFAModelFactory.py:29-32)
kernelMatrixInvChol, kernelMatrixInvPInv = 10000, 10001           # (svGPFAModelFactory.py:25-27)


def buildModelB200(kernels, device=None, process_group=None, shard_mode="auto", indPointsCovRep=indPointsCovChol,
                   kernelMatrixInvMethod=kernelMatrixInvChol):
    """Sibling of ``SVGPFAModelFactory.buildModelPyTorch(kernels=..., indPointsCovRep=...)`` for the in-scope model
    (point process, exponential link, linear embedding, Cholesky Kzz solves; variational covariance as Cholesky
    vectors or rank-1-plus-diagonal; stats/svGPFAModelFactory.py:40-148)."""
    if indPointsCovRep not in (indPointsCovChol, indPointsCovRank1PlusDiag):
        raise ValueError("Invalid indPointsCovRep")
    if kernelMatrixInvMethod not in (kernelMatrixInvChol, kernelMatrixInvPInv):
        raise ValueError("Invalid kernelMatrixInvMethod")
    rep = "rank1_plus_diag" if indPointsCovRep == indPointsCovRank1PlusDiag else "chol"
    inv = "pinv" if kernelMatrixInvMethod == kernelMatrixInvPInv else "chol"
    return B200SVLowerBound(kernels=kernels, device=device, process_group=process_group, shard_mode=shard_mode,
                            ind_points_cov_rep=rep, kzz_inv_method=inv)

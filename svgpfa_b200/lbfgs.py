"""Device-resident, shard-aware L-BFGS for the ECM steps.

The reference maximises every conditional step with ``torch.optim.LBFGS`` (stats/svEM.py:218-294: one fresh
instance per step, ``closure = -eval(); backward()``, ``func_evals`` / ``n_iter`` read back from the state).
``LBFGS`` below has that constructor, ``step(closure)``, ``zero_grad()`` and state, takes the same decisions from
the same quantities (memory update when ``y.s > 1e-10``, ``H = y.s / y.y``, first step ``min(1, 1/|g|_1) lr``, the
strong-Wolfe bracketing / zoom search with cubic interpolation, the four termination tests), and differs in how
the vectors are handled:

* the optimiser state (gradients of the bracket, search direction, line-search origin, the (s, y) history) lives in
  flat device vectors that are allocated once and reused -- no ``clone`` per trial point, no ``torch.cat`` per
  gradient; a trial point is ONE kernel ``x = x0 + t d`` written straight into the packed parameter buffers the
  model's kernels read (``svgpfa_lbfgs_step``), and one small device->host read per closure call brings back
  ``[g.d, max|g|, sum|g|]`` (``svgpfa_lbfgs_stats``) -- torch's loop reads a scalar back for every stored pair;
* the two-loop recursion runs in COEFFICIENT space: the direction is a linear combination of the stored ``s_i``,
  ``y_i`` and the gradient, and the recursion only needs their Gram matrix.  An iteration computes the three new
  rows of that matrix in one pass over the history (``svgpfa_lbfgs_multidot``), runs the recursion on
  ``(2h+1)``-vectors on the host, and forms the direction in a second pass (``svgpfa_lbfgs_combine``, which also
  returns ``g.d`` and ``max|d|``).  At config #5 the E-step vector has 2.2e8 entries: two passes over the history
  against torch's ~10 n-vector sweeps per stored pair;
* **sharding**: with ``process_group`` the parameter vector is the concatenation over ranks of the rank-local
  vectors (trial-sharded leaves: m, cholVecs, Z).  Every reduction the algorithm takes a decision from is global --
  the Gram rows are all-reduced once per iteration, the trial-point statistics once per closure call -- so all
  ranks take identical decisions, call the closure the same number of times (the model's own all-reduce inside
  ``eval()`` stays matched, ``shard_mode="reduce"``) and follow the JOINT trajectory of the single-process
  optimisation: SURVEY.md §8e option (ii) without gathering any gradient.

The vector primitives are hand-written CUDA behind the C ABI (``csrc/lbfgs.cu``); there is no CPU implementation in
the package (the CPU tests inject their own, ``tests/vector_ops_torch.py``).
"""
from __future__ import annotations

import ctypes
import math

import numpy as np
import torch

from . import _cabi

_F64 = torch.float64


class CudaVectorOps:
    """The n-vector primitives of ``include/svgpfa_b200.h`` (svgpfa_lbfgs_*) on one device."""

    def __init__(self, device):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("svgpfa_b200.lbfgs needs a CUDA device: there is no CPU path")
        self._lib = _cabi.lib()
        self._ws = torch.empty(int(self._lib.svgpfa_lbfgs_ws_doubles()), dtype=_F64, device=self.device)
        self.max_vecs = 64                                   # SVGPFA_LBFGS_MAX_VECS

    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @staticmethod
    def _ptrs(tensors):
        return (ctypes.c_void_p * len(tensors))(*[t.data_ptr() for t in tensors])

    def empty(self, n):
        return torch.empty(n, dtype=_F64, device=self.device)

    def multidot(self, vecs, probes):
        """(len(vecs), len(probes)) device tensor of dot products."""
        n, npb = vecs[0].numel(), len(probes)
        out = torch.empty(len(vecs), npb, dtype=_F64, device=self.device)
        pp = self._ptrs(probes)
        with torch.cuda.device(self.device):
            for c0 in range(0, len(vecs), self.max_vecs):
                chunk = vecs[c0:c0 + self.max_vecs]
                _cabi.check(self._lib.svgpfa_lbfgs_multidot(self._ptrs(chunk), len(chunk), pp, npb, n, self._ws.data_ptr(),
                                                            out[c0:].data_ptr(), self._stream()), "lbfgs_multidot")
        return out

    def combine(self, d, vecs, coefs, g):
        """d = sum coefs[i] vecs[i]; returns the device tensor [g.d, max|d|]."""
        out = torch.empty(2, dtype=_F64, device=self.device)
        with torch.cuda.device(self.device):
            for c0 in range(0, len(vecs), self.max_vecs):
                chunk = vecs[c0:c0 + self.max_vecs]
                cf = (ctypes.c_double * len(chunk))(*[float(c) for c in coefs[c0:c0 + self.max_vecs]])
                _cabi.check(self._lib.svgpfa_lbfgs_combine(d.data_ptr(), self._ptrs(chunk), cf, len(chunk), int(c0 > 0),
                                                           g.data_ptr(), d.numel(), self._ws.data_ptr(), out.data_ptr(),
                                                           self._stream()), "lbfgs_combine")
        return out

    def stats(self, a, b=None):
        """Device tensor [a.b, max|a|, sum|a|, max|b|]."""
        out = torch.empty(4, dtype=_F64, device=self.device)
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.svgpfa_lbfgs_stats(a.data_ptr(), b.data_ptr() if b is not None else None, a.numel(),
                                                     self._ws.data_ptr(), out.data_ptr(), self._stream()), "lbfgs_stats")
        return out

    def update(self, s, y, d, t, g, g_prev):
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.svgpfa_lbfgs_update(s.data_ptr(), y.data_ptr(), d.data_ptr(), float(t), g.data_ptr(),
                                                      g_prev.data_ptr(), s.numel(), self._stream()), "lbfgs_update")

    def step(self, x, x0, d, t):
        with torch.cuda.device(self.device):
            _cabi.check(self._lib.svgpfa_lbfgs_step(x.data_ptr(), x0.data_ptr(), d.data_ptr(), float(t), x.numel(),
                                                    self._stream()), "lbfgs_step")


def _cubic_interpolate(x1, f1, g1, x2, f2, g2, bounds=None):
    """Minimiser of the cubic through two points with values and slopes, clipped to ``bounds``
    (the interpolation step of torch.optim.lbfgs's line search, on Python floats)."""
    if bounds is not None:
        lo, hi = bounds
    else:
        lo, hi = (x1, x2) if x1 <= x2 else (x2, x1)
    d1 = g1 + g2 - 3 * (f1 - f2) / (x1 - x2)
    sq = d1 * d1 - g1 * g2
    if sq >= 0:
        d2 = math.sqrt(sq)
        if x1 <= x2:
            pos = x2 - (x2 - x1) * ((g2 + d2 - d1) / (g2 - g1 + 2 * d2))
        else:
            pos = x1 - (x1 - x2) * ((g1 + d2 - d1) / (g1 - g2 + 2 * d2))
        return min(max(pos, lo), hi)
    return (lo + hi) / 2.0


class _Point:
    """A trial point of the line search: step, loss, directional derivative, max|g| and the pool buffer holding g."""
    __slots__ = ("t", "f", "gtd", "gmax", "buf")

    def __init__(self, t, f, gtd, gmax, buf):
        self.t, self.f, self.gtd, self.gmax, self.buf = t, f, gtd, gmax, buf


class LBFGS(torch.optim.Optimizer):
    """Drop-in for ``torch.optim.LBFGS`` on float64 CUDA leaves (see the module docstring).

    ``process_group``: the leaves are this rank's shard of a larger vector; decisions are taken from globally reduced
    quantities and the closure must return the GLOBAL loss (``B200SVLowerBound(shard_mode="reduce")``).
    ``ops``: the vector primitives (default: the CUDA library; tests inject a torch implementation)."""

    def __init__(self, params, lr=1, max_iter=20, max_eval=None, tolerance_grad=1e-7, tolerance_change=1e-9,
                 history_size=100, line_search_fn=None, *, process_group=None, ops=None):
        if not 0.0 <= lr:
            raise ValueError(f"Invalid learning rate: {lr}")
        if line_search_fn not in (None, "strong_wolfe"):
            raise RuntimeError("only 'strong_wolfe' is supported")
        if max_eval is None:
            max_eval = max_iter * 5 // 4
        defaults = dict(lr=lr, max_iter=max_iter, max_eval=max_eval, tolerance_grad=tolerance_grad,
                        tolerance_change=tolerance_change, history_size=history_size, line_search_fn=line_search_fn)
        super().__init__(params, defaults)
        if len(self.param_groups) != 1:
            raise ValueError("LBFGS doesn't support per-parameter options (parameter groups)")
        self._params = self.param_groups[0]["params"]
        for p in self._params:
            if p.dtype != _F64 or not p.is_contiguous():
                raise ValueError("svgpfa_b200.lbfgs.LBFGS needs contiguous float64 parameters")
        self._pg = process_group
        self._ops = ops if ops is not None else CudaVectorOps(self._params[0].device)
        self._n = sum(p.numel() for p in self._params)
        self._offsets = np.concatenate([[0], np.cumsum([p.numel() for p in self._params])]).tolist()
        self._segments = self._find_segments()
        self._bases = self._find_bases()
        # flat vectors, allocated on first use
        self._d = self._x0 = self._g_prev = None
        self._pool = []                                      # gradient buffers of the line search (4)
        self._g = None                                       # index into the pool of the current gradient
        self._slots, self._order, self._free = [], [], []    # (s, y) pairs: storage, oldest -> newest, recycled
        self._ro = []
        self._H = 1.0
        self._SS = self._SY = self._YY = np.zeros((0, 0))
        self._t = None
        self._prev_loss = None
        self.host_reads = 0                                  # device -> host reads taken so far (measurement)

    # ------------------------------------------------------------------ flat views of the leaves
    def _find_segments(self):
        """Runs of leaves that are adjacent in memory (the model's leaves are views of one packed buffer per group,
        laid out in getter order) become ONE segment: a trial point is one kernel per segment, not one per leaf."""
        segs = []
        for i, p in enumerate(self._params):
            ptr, numel = p.data_ptr(), p.numel()
            if segs and segs[-1]["end"] == ptr and segs[-1]["storage"] == p.untyped_storage().data_ptr():
                segs[-1]["end"] = ptr + 8 * numel
                segs[-1]["numel"] += numel
            else:
                segs.append(dict(first=i, off=self._offsets[i], numel=numel, end=ptr + 8 * numel,
                                 storage=p.untyped_storage().data_ptr()))
        out = []
        for s in segs:
            p = self._params[s["first"]]
            view = torch.empty(0, dtype=_F64, device=p.device).set_(p.untyped_storage(), p.storage_offset(), (s["numel"],), (1,))
            out.append((view, s["off"], s["numel"]))
        return out

    def _find_bases(self):
        """One tensor per version counter: the kernels write through raw pointers, so the counters the model keys its
        caches on (Kzz / Cholesky factors, spike statistics) are bumped by an empty in-place operation afterwards."""
        seen, bases = set(), []
        for p in self._params:
            b = p._base if p._base is not None else p
            if id(b) not in seen:
                seen.add(id(b))
                bases.append(b)
        return bases

    def _bump_versions(self):
        for b in self._bases:
            b.detach().view(-1)[:0].zero_()

    def _ensure_buffers(self):
        if self._d is None:
            e = self._ops.empty
            self._d, self._x0, self._g_prev = e(self._n), e(self._n), e(self._n)
            self._pool = [e(self._n) for _ in range(4)]

    def _gather_params(self, out):
        for view, off, numel in self._segments:
            out[off:off + numel].copy_(view)

    def _set_params(self, x0, d, t):
        """params = x0 + t d."""
        for view, off, numel in self._segments:
            self._ops.step(view, x0[off:off + numel], d[off:off + numel], t)
        self._bump_versions()

    def _gather_grad(self, out):
        dst, src = [], []
        for i, p in enumerate(self._params):
            o = out[self._offsets[i]:self._offsets[i + 1]]
            if p.grad is None:
                o.zero_()
            else:
                dst.append(o)
                src.append(p.grad.reshape(-1))
        if dst:
            torch._foreach_copy_(dst, src)

    # ------------------------------------------------------------------ reductions (global under sharding)
    def _read(self, dev_tensor):
        self.host_reads += 1
        return dev_tensor.cpu().numpy()

    def _point_stats(self, loss, g, d):
        """[loss, g.d, max|g|, sum|g|, max|d|] in one device->host read."""
        st = self._ops.stats(g, d)
        if self._pg is None:
            v = self._read(torch.cat([st, loss.detach().reshape(1).to(st.dtype)]))
            return float(v[4]), float(v[0]), float(v[1]), float(v[2]), float(v[3])
        import torch.distributed as dist
        W = dist.get_world_size(self._pg)
        allst = torch.empty(W * 4, dtype=st.dtype, device=st.device)
        dist.all_gather_into_tensor(allst, st, group=self._pg)
        v = self._read(torch.cat([allst, loss.detach().reshape(1).to(st.dtype)]))
        a = v[:-1].reshape(W, 4)                             # summed / maximised in rank order: identical on every rank
        return float(v[-1]), float(a[:, 0].sum()), float(a[:, 1].max()), float(a[:, 2].sum()), float(a[:, 3].max())

    def _reduce_sum(self, dev_tensor):
        if self._pg is not None:
            import torch.distributed as dist
            dist.all_reduce(dev_tensor, op=dist.ReduceOp.SUM, group=self._pg)
        return self._read(dev_tensor)

    def _reduce_gd(self, out2):
        """[g.d, max|d|] of the combine kernel."""
        if self._pg is None:
            v = self._read(out2)
            return float(v[0]), float(v[1])
        import torch.distributed as dist
        W = dist.get_world_size(self._pg)
        allv = torch.empty(W * 2, dtype=out2.dtype, device=out2.device)
        dist.all_gather_into_tensor(allv, out2, group=self._pg)
        a = self._read(allv).reshape(W, 2)
        return float(a[:, 0].sum()), float(a[:, 1].max())

    # ------------------------------------------------------------------ history and direction
    def _new_slot(self):
        if self._free:
            return self._free.pop()
        self._slots.append((self._ops.empty(self._n), self._ops.empty(self._n)))
        return len(self._slots) - 1

    def _update_memory_and_direction(self, t, history_size):
        """torch/optim/lbfgs.py "do lbfgs update (update memory)" + two-loop recursion, in coefficient space.
        Returns (g.d, max|d|) of the new direction, which is left in ``self._d``."""
        ops, g = self._ops, self._pool[self._g]
        cand = self._new_slot()
        s_c, y_c = self._slots[cand]
        ops.update(s_c, y_c, self._d, t, g, self._g_prev)                  # s = t d, y = g - g_prev, g_prev = g
        h = len(self._order)
        S = [self._slots[i][0] for i in self._order] + [s_c]
        Y = [self._slots[i][1] for i in self._order] + [y_c]
        R = self._reduce_sum(ops.multidot(S + Y + [g], [s_c, y_c, g]))      # (2 (h + 1) + 1, 3)
        h1 = h + 1
        ys, yy = float(R[h, 1]), float(R[h1 + h, 1])
        Sg, Yg, gg = R[:h1, 2].copy(), R[h1:2 * h1, 2].copy(), float(R[2 * h1, 2])
        if ys > 1e-10:
            SS, SY, YY = np.zeros((h1, h1)), np.zeros((h1, h1)), np.zeros((h1, h1))
            SS[:h, :h], SY[:h, :h], YY[:h, :h] = self._SS, self._SY, self._YY
            SS[h, :], SS[:, h] = R[:h1, 0], R[:h1, 0]
            YY[h, :], YY[:, h] = R[h1:2 * h1, 1], R[h1:2 * h1, 1]
            SY[:h1, h] = R[:h1, 1]                                           # s_i . y_c
            SY[h, :h1] = R[h1:2 * h1, 0]                                     # s_c . y_j
            SY[h, h] = ys
            self._order.append(cand)
            self._ro.append(1.0 / ys)
            self._H = ys / yy
            if len(self._order) > history_size:                              # shift the history by one
                self._free.append(self._order.pop(0))
                self._ro.pop(0)
                SS, SY, YY, Sg, Yg = SS[1:, 1:], SY[1:, 1:], YY[1:, 1:], Sg[1:], Yg[1:]
            self._SS, self._SY, self._YY = SS, SY, YY
        else:
            self._free.append(cand)
            Sg, Yg = Sg[:h], Yg[:h]
        h = len(self._order)
        SS, SY, YY, ro = self._SS, self._SY, self._YY, self._ro
        dS, dY, dg = np.zeros(h), np.zeros(h), -1.0                          # q = -g
        al = np.zeros(h)
        for i in range(h - 1, -1, -1):
            al[i] = (SS[i] @ dS + SY[i] @ dY + Sg[i] * dg) * ro[i]           # (s_i . q) ro_i
            dY[i] -= al[i]                                                   # q -= al_i y_i
        dS *= self._H
        dY *= self._H
        dg *= self._H
        for i in range(h):
            be = (SY[:, i] @ dS + YY[i] @ dY + Yg[i] * dg) * ro[i]           # (y_i . r) ro_i
            dS[i] += al[i] - be                                              # r += (al_i - be_i) s_i
        vecs = [self._slots[i][0] for i in self._order] + [self._slots[i][1] for i in self._order] + [g]
        return self._reduce_gd(ops.combine(self._d, vecs, list(dS) + list(dY) + [dg], g))

    # ------------------------------------------------------------------ line search
    def _free_buffer(self, *in_use):
        used = {b for b in in_use if b is not None}
        for i in range(len(self._pool)):
            if i not in used:
                return i
        raise AssertionError("line-search gradient pool exhausted")

    def _evaluate(self, closure, t, buf):
        """Loss and gradient at x0 + t d (gradient into pool buffer ``buf``)."""
        self._set_params(self._x0, self._d, t)
        loss = closure()
        g = self._pool[buf]
        self._gather_grad(g)
        f, gtd, gmax, _, _ = self._point_stats(loss, g, self._d)
        return _Point(t, f, gtd, gmax, buf)

    def _strong_wolfe(self, closure, t, f, gtd, d_norm, c1=1e-4, c2=0.9, tolerance_change=1e-9, max_ls=25):
        """The bracketing / zoom search of torch.optim.lbfgs (ported there from minFunc's lswolfe) with the gradients of
        the bracket kept in a pool of flat buffers.  Returns (point, function evaluations)."""
        start = _Point(0.0, f, gtd, None, self._g)
        new = self._evaluate(closure, t, self._free_buffer(start.buf))
        evals = 1
        prev = start
        bracket, done, ls_iter = None, False, 0
        while ls_iter < max_ls:
            if new.f > (f + c1 * new.t * gtd) or (ls_iter > 1 and new.f >= prev.f):
                bracket = [prev, new]
                break
            if abs(new.gtd) <= -c2 * gtd:
                bracket, done = [new], True
                break
            if new.gtd >= 0:
                bracket = [prev, new]
                break
            min_step = new.t + 0.01 * (new.t - prev.t)
            max_step = new.t * 10
            t_next = _cubic_interpolate(prev.t, prev.f, prev.gtd, new.t, new.f, new.gtd, bounds=(min_step, max_step))
            older, prev = prev, new
            new = self._evaluate(closure, t_next, self._free_buffer(start.buf, prev.buf))
            del older
            evals += 1
            ls_iter += 1
        if ls_iter == max_ls:
            bracket = [start, new]
        insuf_progress = False
        low, high = (0, 1) if bracket[0].f <= bracket[-1].f else (1, 0)
        while not done and ls_iter < max_ls:
            if abs(bracket[1].t - bracket[0].t) * d_norm < tolerance_change:
                break
            t_new = _cubic_interpolate(bracket[0].t, bracket[0].f, bracket[0].gtd, bracket[1].t, bracket[1].f, bracket[1].gtd)
            bmax, bmin = max(bracket[0].t, bracket[1].t), min(bracket[0].t, bracket[1].t)
            eps = 0.1 * (bmax - bmin)
            if min(bmax - t_new, t_new - bmin) < eps:
                if insuf_progress or t_new >= bmax or t_new <= bmin:
                    t_new = bmax - eps if abs(t_new - bmax) < abs(t_new - bmin) else bmin + eps
                    insuf_progress = False
                else:
                    insuf_progress = True
            else:
                insuf_progress = False
            new = self._evaluate(closure, t_new, self._free_buffer(start.buf, bracket[0].buf, bracket[1].buf))
            evals += 1
            ls_iter += 1
            if new.f > (f + c1 * new.t * gtd) or new.f >= bracket[low].f:
                bracket[high] = new
                low, high = (0, 1) if bracket[0].f <= bracket[1].f else (1, 0)
            else:
                if abs(new.gtd) <= -c2 * gtd:
                    done = True
                elif new.gtd * (bracket[high].t - bracket[low].t) >= 0:
                    bracket[high] = bracket[low]
                bracket[low] = new
        return bracket[low] if len(bracket) > 1 else bracket[0], evals

    # ------------------------------------------------------------------ the step
    @torch.no_grad()
    def step(self, closure):
        closure = torch.enable_grad()(closure)
        group = self.param_groups[0]
        lr = float(group["lr"])
        max_iter, max_eval = group["max_iter"], group["max_eval"]
        tolerance_grad, tolerance_change = group["tolerance_grad"], group["tolerance_change"]
        line_search_fn, history_size = group["line_search_fn"], group["history_size"]
        state = self.state[self._params[0]]
        state.setdefault("func_evals", 0)
        state.setdefault("n_iter", 0)
        self._ensure_buffers()

        orig_loss = closure()
        current_evals = 1
        state["func_evals"] += 1
        if self._g is None:
            self._g = 0
        self._gather_grad(self._pool[self._g])
        loss, _, gmax, gsum, _ = self._point_stats(orig_loss, self._pool[self._g], None)
        if gmax <= tolerance_grad:
            return orig_loss

        t = self._t
        n_iter = 0
        while n_iter < max_iter:
            n_iter += 1
            state["n_iter"] += 1
            g = self._pool[self._g]
            if state["n_iter"] == 1:
                self._order, self._ro, self._H = [], [], 1.0
                self._free = list(range(len(self._slots)))
                self._SS = self._SY = self._YY = np.zeros((0, 0))
                gtd, d_norm = self._reduce_gd(self._ops.combine(self._d, [g], [-1.0], g))      # d = -g
                self._g_prev.copy_(g)
            else:
                gtd, d_norm = self._update_memory_and_direction(t, history_size)
            prev_loss = loss
            if state["n_iter"] == 1:
                t = min(1.0, 1.0 / gsum) * lr
            else:
                t = lr
            if gtd > -tolerance_change:
                break
            ls_func_evals = 0
            self._gather_params(self._x0)
            if line_search_fn is not None:
                point, ls_func_evals = self._strong_wolfe(closure, t, loss, gtd, d_norm,
                                                          max_ls=max_eval - current_evals)
                loss, t, self._g, gmax = point.f, point.t, point.buf, (point.gmax if point.gmax is not None else gmax)
                self._set_params(self._x0, self._d, t)
                opt_cond = gmax <= tolerance_grad
            else:
                opt_cond = False
                if n_iter != max_iter:
                    point = self._evaluate(closure, t, self._g)                 # the old gradient lives on in g_prev
                    loss, gmax, opt_cond = point.f, point.gmax, point.gmax <= tolerance_grad
                    ls_func_evals = 1
                else:
                    self._set_params(self._x0, self._d, t)
            current_evals += ls_func_evals
            state["func_evals"] += ls_func_evals
            if n_iter == max_iter:
                break
            if current_evals >= max_eval:
                break
            if opt_cond:
                break
            if abs(t) * d_norm <= tolerance_change:
                break
            if abs(loss - prev_loss) < tolerance_change:
                break
        self._t = t
        self._prev_loss = loss
        return orig_loss


class patched_torch_lbfgs:
    """``with patched_torch_lbfgs(): SVEM_PyTorch().maximize(model, ...)`` -- the unmodified reference constructs
    ``torch.optim.LBFGS`` by name (stats/svEM.py:221,229,243,262); inside the block that name is this module's
    optimiser (extra keyword arguments, e.g. ``process_group``, are bound here)."""

    def __init__(self, **extra):
        self._extra = extra
        self._saved = None

    def __enter__(self):
        self._saved = torch.optim.LBFGS
        if self._extra:
            extra = self._extra

            class _Bound(LBFGS):
                def __init__(self, params, **kw):
                    super().__init__(params, **dict(extra, **kw))
            torch.optim.LBFGS = _Bound
        else:
            torch.optim.LBFGS = LBFGS
        return self

    def __exit__(self, *exc):
        torch.optim.LBFGS = self._saved
        return False

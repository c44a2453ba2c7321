"""torch.autograd glue between the reference-shaped Python functions and the C ABI.

PyTorch supplies device memory, streams and the tape only; every arithmetic step is a libkccot
kernel.  Tensors must be CUDA fp32 — anything else raises ValueError (no fallback).
"""
import ctypes

import torch

from . import _lib

# default kernel-path flag; tests flip it through `set_path` to cross-check tcgen05 vs CUDA cores
_PATH = {"flags": _lib.PATH_AUTO}


def set_path(name):
    """'auto' | 'simt' | 'tcgen05' — which kernel family the cost GEMMs use (tests only)."""
    _PATH["flags"] = {"auto": _lib.PATH_AUTO, "simt": _lib.PATH_SIMT, "tcgen05": _lib.PATH_TCGEN05}[name]


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream(dev):
    # raw handle of the current stream of `dev` (torch.cuda.current_stream(dev).cuda_stream costs ~10 us per call)
    return ctypes.c_void_p(torch._C._cuda_getCurrentRawStream(dev.index if dev.index is not None else torch.cuda.current_device()))


class _on_device:
    """`with torch.cuda.device(dev)` only when `dev` is not already the current device (the common case costs one
    query instead of two device switches)."""
    __slots__ = ("dev", "ctx")

    def __init__(self, dev):
        self.dev = dev
        self.ctx = None

    def __enter__(self):
        idx = self.dev.index
        if idx is not None and idx != torch._C._cuda_getDevice():
            self.ctx = torch.cuda.device(self.dev)
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)
        return False


def _check(t, name, ndim=None):
    if not isinstance(t, torch.Tensor):
        raise ValueError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise ValueError(f"{name}: expected a CUDA tensor (kccotgan_b200 has no CPU path), got {t.device}")
    if t.dtype != torch.float32:
        raise ValueError(f"{name}: expected float32, got {t.dtype}")
    if ndim is not None and t.dim() != ndim:
        raise ValueError(f"{name}: expected {ndim} dimensions, got shape {tuple(t.shape)}")
    return t.contiguous()


def _ws(nbytes, dev):
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=dev)


def _flat_rows(x, name):
    x = _check(x, name)
    if x.dim() < 2:
        raise ValueError(f"{name}: expected [batch, ...], got shape {tuple(x.shape)}")
    return x.reshape(x.shape[0], -1)


def _same(a, b):
    return a is b or (a.data_ptr() == b.data_ptr() and a.shape == b.shape)


# ------------------------------------------------------------------------------------------------
# cost matrix: cost_xy / modified_cost / bi_causal_modified_cost
# ------------------------------------------------------------------------------------------------
class CostFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, h1, M1, h2, M2, s):
        X, Y = _flat_rows(x, "x"), _flat_rows(y, "y")
        if X.shape[1] != Y.shape[1]:
            raise ValueError(f"x and y disagree on the feature size: {tuple(x.shape)} vs {tuple(y.shape)}")
        same = _same(X, Y)
        if same:
            Y = X
        Bx, K = X.shape
        By = Y.shape[0]
        pairs = []
        T = J = 0
        for h, M, tag in ((h1, M1, "1"), (h2, M2, "2")):
            if h is None:
                pairs.append((None, None))
                continue
            h, M = _check(h, "h" + tag, 3), _check(M, "M" + tag, 3)
            if h.shape[0] != Bx or M.shape[0] != By or h.shape[1:] != M.shape[1:]:
                raise ValueError(f"h{tag}/M{tag} shapes {tuple(h.shape)}/{tuple(M.shape)} do not match x/y batches "
                                 f"{Bx}/{By}")
            if T and (h.shape[1], h.shape[2]) != (T, J):
                raise ValueError("both martingale pairs must share [T, J]")
            T, J = h.shape[1], h.shape[2]
            pairs.append((h, M))
        dev = X.device
        C = torch.empty((Bx, By), dtype=torch.float32, device=dev)
        lib = _lib.load()
        nb = lib.kccot_cost_workspace_bytes(1, Bx, By, K)
        ws = _ws(nb, dev)
        with _on_device(dev):
            _lib.call("kccot_cost_fwd", _ptr(X), _ptr(Y), 1, Bx, By, K, _ptr(pairs[0][0]), _ptr(pairs[0][1]),
                      _ptr(pairs[1][0]), _ptr(pairs[1][1]), T, J, float(s), _ptr(C), _ptr(ws), ws.numel(),
                      _PATH["flags"], _stream(dev))
        ctx.save_for_backward(X, Y, *[t for p in pairs for t in p if t is not None])
        ctx.meta = (x.shape, y.shape, same, [p[0] is not None for p in pairs], T, J, float(s))
        return C

    @staticmethod
    def backward(ctx, gC):
        xshape, yshape, same, has, T, J, s = ctx.meta
        saved = list(ctx.saved_tensors)
        X, Y = saved[0], saved[1]
        rest = saved[2:]
        pairs = []
        for hflag in has:
            pairs.append((rest.pop(0), rest.pop(0)) if hflag else (None, None))
        gC = gC.contiguous().float()
        dev = X.device
        Bx, K = X.shape
        By = Y.shape[0]
        need = ctx.needs_input_grad
        gx = gy = None
        with _on_device(dev):
            st = _stream(dev)
            if need[0] or need[1]:
                gx = torch.empty_like(X) if need[0] else None
                gy = torch.empty_like(Y) if need[1] else None
                ws = _ws(_lib.load().kccot_cost_bwd_workspace_bytes(1, Bx, By, K), dev)
                _lib.call("kccot_cost_bwd", _ptr(gC), _ptr(X), _ptr(Y), 1, Bx, By, K, s, _ptr(gx), _ptr(gy),
                          _ptr(ws), ws.numel(), _PATH["flags"] if not same else _lib.PATH_SIMT, st)
            gout = [None, None, None, None]
            for q, (h, M) in enumerate(pairs):
                if h is None or not (need[2 + 2 * q] or need[3 + 2 * q]):
                    continue
                gh = torch.empty_like(h) if need[2 + 2 * q] else None
                gM = torch.empty_like(M) if need[3 + 2 * q] else None
                _lib.call("kccot_martingale_bwd", _ptr(gC), _ptr(h), _ptr(M), 1, Bx, By, T, J, s, _ptr(gh), _ptr(gM),
                          0, st)
                gout[2 * q], gout[2 * q + 1] = gh, gM
        gx = gx.reshape(xshape) if gx is not None else None
        gy = gy.reshape(yshape) if gy is not None else None
        return (gx, gy, *gout, None)


# ------------------------------------------------------------------------------------------------
# Sinkhorn solve on a given cost: C [n,B,B] -> cost [n]
# ------------------------------------------------------------------------------------------------
class SinkhornFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, C, eps, L, Lmin, thresh, exit_on_index):
        C = _check(C, "C", 3)
        n, B, B2 = C.shape
        if B != B2:
            raise ValueError(f"Sinkhorn needs square cost matrices, got {tuple(C.shape)}")
        dev = C.device
        L = int(L)
        uh = torch.empty((n, L + 1, B), dtype=torch.float32, device=dev)
        vh = torch.empty((n, L + 1, B), dtype=torch.float32, device=dev)
        nits = torch.empty((n,), dtype=torch.int32, device=dev)
        cost = torch.empty((n,), dtype=torch.float32, device=dev)
        ws = _ws(_lib.load().kccot_sinkhorn_workspace_bytes(n, B, L), dev)
        with _on_device(dev):
            _lib.call("kccot_sinkhorn_fwd", _ptr(C), n, B, float(eps), L, int(Lmin), float(thresh),
                      int(bool(exit_on_index)), _ptr(uh), _ptr(vh), _ptr(nits), _ptr(cost), _ptr(ws), ws.numel(),
                      _stream(dev))
        ctx.save_for_backward(C, uh, vh, nits)
        ctx.meta = (float(eps), L)
        ctx.mark_non_differentiable(nits)
        return cost, nits

    @staticmethod
    def backward(ctx, gcost, _gnits):
        C, uh, vh, nits = ctx.saved_tensors
        eps, L = ctx.meta
        n, B, _ = C.shape
        dev = C.device
        gcost = gcost.contiguous().float()
        Cbar = torch.empty_like(C)
        ws = _ws(_lib.load().kccot_sinkhorn_workspace_bytes(n, B, L), dev)
        with _on_device(dev):
            _lib.call("kccot_sinkhorn_bwd", _ptr(C), n, B, eps, L, _ptr(uh), _ptr(vh), _ptr(nits), _ptr(gcost),
                      _ptr(Cbar), _ptr(ws), ws.numel(), _stream(dev))
        return Cbar, None, None, None, None, None


# ------------------------------------------------------------------------------------------------
# fused mixed loss: 2*S(real,fake) - S(real,real) - S(fake,fake), one pass over the videos
# ------------------------------------------------------------------------------------------------
_SIZES = {}


def _mixed_sizes(B, K, L, nprob=1):
    key = (B, K, L, nprob)
    if key not in _SIZES:
        lib = _lib.load()
        _SIZES[key] = (int(lib.kccot_mixed_loss_saved_bytes(nprob, B, L)),
                       int(lib.kccot_mixed_loss_workspace_bytes(nprob, B, K, L)))
    return _SIZES[key]


class MixedLossFn(torch.autograd.Function):
    """One C-ABI call per direction (kccot_mixed_loss_fwd / _bwd)."""

    @staticmethod
    def forward(ctx, real, fake, h_fake, m_real, h_real, m_fake, s, eps, L, shared=(0, 0)):
        """shared = (period, len) in flattened columns: the shared-context hint of kccot_mixed_loss_*_ctx."""
        R, F = _flat_rows(real, "f_real"), _flat_rows(fake, "f_fake")
        if R.shape != F.shape:
            raise ValueError(f"f_real and f_fake must have the same shape, got {tuple(real.shape)} vs "
                             f"{tuple(fake.shape)}")
        B, K = R.shape
        hs = [_check(t, n, 3) for t, n in ((h_fake, "h_fake"), (m_real, "m_real"), (h_real, "h_real"),
                                           (m_fake, "m_fake"))]
        T, J = hs[0].shape[1], hs[0].shape[2]
        for t, n in zip(hs, ("h_fake", "m_real", "h_real", "m_fake")):
            if tuple(t.shape) != (B, T, J):
                raise ValueError(f"{n}: expected shape {(B, T, J)}, got {tuple(t.shape)}")
        if T < 2:
            raise ValueError(f"the martingale term needs at least 2 time steps, got T={T}")
        dev = R.device
        L = int(L)
        saved_bytes, ws_bytes = _mixed_sizes(B, K, L)
        saved = torch.empty(saved_bytes, dtype=torch.uint8, device=dev)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        out = torch.empty(4, dtype=torch.float32, device=dev)          # loss | xy, xx, yy
        with _on_device(dev):
            _lib.call("kccot_mixed_loss_fwd_ctx", _ptr(R), _ptr(F), 1, B, K, _ptr(hs[0]), _ptr(hs[1]), _ptr(hs[2]),
                      _ptr(hs[3]), T, J, float(s), float(eps), L, _ptr(saved), ctypes.c_void_p(out.data_ptr()),
                      ctypes.c_void_p(out.data_ptr() + 4), _ptr(ws), ws_bytes, _PATH["flags"], _stream(dev),
                      int(shared[0]), int(shared[1]))
        ctx.save_for_backward(R, F, *hs, saved)
        ctx.meta = (real.shape, fake.shape, float(s), float(eps), L, (int(shared[0]), int(shared[1])))
        loss, terms = out[0], out[1:]
        ctx.mark_non_differentiable(terms)
        ctx.set_materialize_grads(False)          # no zero-fill kernel for the unused gradient of `terms`
        return loss, terms

    @staticmethod
    def backward(ctx, gloss, _gterms):
        R, F, h_fake, m_real, h_real, m_fake, saved = ctx.saved_tensors
        rshape, fshape, s, eps, L, shared = ctx.meta
        B, K = R.shape
        T, J = h_fake.shape[1], h_fake.shape[2]
        dev = R.device
        need = ctx.needs_input_grad
        if gloss is None:                         # loss itself unused downstream
            return (None,) * 10
        gloss = gloss.reshape(1)
        if gloss.dtype != torch.float32 or not gloss.is_contiguous():
            gloss = gloss.float().contiguous()
        g_real = torch.empty_like(R) if need[0] else None
        g_fake = torch.empty_like(F) if need[1] else None
        gh_fake = torch.empty_like(h_fake) if need[2] else None
        gm_real = torch.empty_like(m_real) if need[3] else None
        gh_real = torch.empty_like(h_real) if need[4] else None
        gm_fake = torch.empty_like(m_fake) if need[5] else None
        _, ws_bytes = _mixed_sizes(B, K, L)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with _on_device(dev):
            _lib.call("kccot_mixed_loss_bwd_ctx", _ptr(gloss), _ptr(R), _ptr(F), 1, B, K, _ptr(h_fake), _ptr(m_real),
                      _ptr(h_real), _ptr(m_fake), T, J, s, eps, L, _ptr(saved), _ptr(g_real), _ptr(g_fake),
                      _ptr(gh_fake), _ptr(gm_real), _ptr(gh_real), _ptr(gm_fake), _ptr(ws), ws_bytes, _PATH["flags"],
                      _stream(dev), shared[0], shared[1])
        g_real = g_real.reshape(rshape) if g_real is not None else None
        g_fake = g_fake.reshape(fshape) if g_fake is not None else None
        return g_real, g_fake, gh_fake, gm_real, gh_real, gm_fake, None, None, None, None


# ------------------------------------------------------------------------------------------------
# martingale penalty
# ------------------------------------------------------------------------------------------------
class MixedLossBatchedFn(torch.autograd.Function):
    """`nprob` independent problems in one C-ABI call per direction (BASELINE config 4): every tensor carries a
    leading problem axis, the result is one loss per problem."""

    @staticmethod
    def forward(ctx, real, fake, h_fake, m_real, h_real, m_fake, s, eps, L):
        real, fake = _check(real, "f_real"), _check(fake, "f_fake")
        if real.dim() < 3 or real.shape != fake.shape:
            raise ValueError(f"f_real / f_fake: expected equal shapes [nprob, B, ...], got {tuple(real.shape)} vs "
                             f"{tuple(fake.shape)}")
        P, B = real.shape[0], real.shape[1]
        R, F = real.reshape(P, B, -1), fake.reshape(P, B, -1)
        K = R.shape[2]
        hs = [_check(t, n, 4) for t, n in ((h_fake, "h_fake"), (m_real, "m_real"), (h_real, "h_real"),
                                           (m_fake, "m_fake"))]
        T, J = hs[0].shape[2], hs[0].shape[3]
        for t, n in zip(hs, ("h_fake", "m_real", "h_real", "m_fake")):
            if tuple(t.shape) != (P, B, T, J):
                raise ValueError(f"{n}: expected shape {(P, B, T, J)}, got {tuple(t.shape)}")
        if T < 2:
            raise ValueError(f"the martingale term needs at least 2 time steps, got T={T}")
        dev = R.device
        L = int(L)
        saved_bytes, ws_bytes = _mixed_sizes(B, K, L, P)
        saved = torch.empty(saved_bytes, dtype=torch.uint8, device=dev)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        loss = torch.empty(P, dtype=torch.float32, device=dev)
        terms = torch.empty((P, 3), dtype=torch.float32, device=dev)
        with _on_device(dev):
            _lib.call("kccot_mixed_loss_fwd", _ptr(R), _ptr(F), P, B, K, _ptr(hs[0]), _ptr(hs[1]), _ptr(hs[2]),
                      _ptr(hs[3]), T, J, float(s), float(eps), L, _ptr(saved), _ptr(loss), _ptr(terms), _ptr(ws),
                      ws_bytes, _PATH["flags"], _stream(dev))
        ctx.save_for_backward(R, F, *hs, saved)
        ctx.meta = (real.shape, fake.shape, float(s), float(eps), L)
        ctx.mark_non_differentiable(terms)
        ctx.set_materialize_grads(False)
        return loss, terms

    @staticmethod
    def backward(ctx, gloss, _gterms):
        R, F, h_fake, m_real, h_real, m_fake, saved = ctx.saved_tensors
        rshape, fshape, s, eps, L = ctx.meta
        P, B, K = R.shape
        T, J = h_fake.shape[2], h_fake.shape[3]
        dev = R.device
        need = ctx.needs_input_grad
        if gloss is None:
            return (None,) * 9
        gloss = gloss.reshape(P).float().contiguous()
        outs = [torch.empty_like(t) if need[i] else None for i, t in enumerate((R, F, h_fake, m_real, h_real, m_fake))]
        _, ws_bytes = _mixed_sizes(B, K, L, P)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with _on_device(dev):
            _lib.call("kccot_mixed_loss_bwd", _ptr(gloss), _ptr(R), _ptr(F), P, B, K, _ptr(h_fake), _ptr(m_real),
                      _ptr(h_real), _ptr(m_fake), T, J, s, eps, L, _ptr(saved), *[_ptr(o) for o in outs], _ptr(ws),
                      ws_bytes, _PATH["flags"], _stream(dev))
        if outs[0] is not None:
            outs[0] = outs[0].reshape(rshape)
        if outs[1] is not None:
            outs[1] = outs[1].reshape(fshape)
        return (*outs, None, None, None)


class MartingalePenaltyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, M, reg_lam, s):
        M = _check(M, "M", 3)
        B, T, J = M.shape
        if T < 2:
            raise ValueError(f"M needs at least 2 time steps, got shape {tuple(M.shape)}")
        dev = M.device
        pm = torch.empty((), dtype=torch.float32, device=dev)
        stats = torch.empty((2 * J + (T - 1) * J,), dtype=torch.float32, device=dev)
        with _on_device(dev):
            _lib.call("kccot_pm_fwd", _ptr(M), B, T, J, float(reg_lam), float(s), _ptr(pm), _ptr(stats), _stream(dev))
        ctx.save_for_backward(M, stats)
        ctx.meta = (float(reg_lam), float(s))
        return pm

    @staticmethod
    def backward(ctx, gpm):
        M, stats = ctx.saved_tensors
        lam, s = ctx.meta
        B, T, J = M.shape
        gM = torch.empty_like(M)
        gpm = gpm.reshape(1).contiguous().float()
        with torch.cuda.device(M.device):
            _lib.call("kccot_pm_bwd", _ptr(M), B, T, J, lam, s, _ptr(stats), _ptr(gpm), _ptr(gM), _stream(M.device))
        return gM, None, None


# ------------------------------------------------------------------------------------------------
# Gaussian smoothing
# ------------------------------------------------------------------------------------------------
def _taps(w):
    """Host weights -> (ctypes float array, radius); None -> (None, 0)."""
    if w is None:
        return None, 0
    arr = (ctypes.c_float * len(w))(*[float(v) for v in w])
    return arr, (len(w) - 1) // 2


class SmoothFn(torch.autograd.Function):
    """mode 1: `taps_t` filters T; mode 3: `taps_s` filters H, T and W.  The weights are host sequences (they become
    kernel arguments: nothing is copied to the device, so an annealed sigma costs nothing per step)."""

    @staticmethod
    def forward(ctx, x, mode, taps_t, taps_s):
        x = _check(x, "inputs", 5)
        B, H, T, W, C = x.shape
        dev = x.device
        out = torch.empty_like(x)
        maxval = torch.empty((1,), dtype=torch.float32, device=dev)
        ws = _ws(_lib.load().kccot_smooth_workspace_bytes(mode, B, H, T, W, C), dev)
        at, rt = _taps(taps_t)
        as_, rs = _taps(taps_s)
        with _on_device(dev):
            _lib.call("kccot_smooth_fwd", mode, _ptr(x), B, H, T, W, C, at, rt, as_, rs, _ptr(out), _ptr(maxval), _ptr(ws),
                      ws.numel(), _stream(dev))
        ctx.save_for_backward(out, maxval)
        ctx.mode, ctx.taps = mode, (taps_t, taps_s)
        return out

    @staticmethod
    def backward(ctx, gout):
        out, maxval = ctx.saved_tensors
        B, H, T, W, C = out.shape
        dev = out.device
        gout = gout.contiguous().float()
        gx = torch.empty_like(out)
        ws = _ws(_lib.load().kccot_smooth_workspace_bytes(ctx.mode, B, H, T, W, C), dev)
        at, rt = _taps(ctx.taps[0])
        as_, rs = _taps(ctx.taps[1])
        with _on_device(dev):
            _lib.call("kccot_smooth_bwd", ctx.mode, _ptr(gout), _ptr(out), _ptr(maxval), B, H, T, W, C, at, rt, as_, rs,
                      _ptr(gx), _ptr(ws), ws.numel(), _stream(dev))
        return gx, None, None, None

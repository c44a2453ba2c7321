"""fp64 numpy restatement of the causal-OT loss path, with hand-derived backward passes.

TEST INFRASTRUCTURE ONLY — imported by `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU
legs; never by `kccotgan_b200`.  Every function cites the reference lines it restates
(paths are into neuripss2020/kccotgan).  Forward passes follow the reference's formulation; the
backward passes are the analytic reverse-mode recurrences of SURVEY.md Appendix A, written
independently of any autograd so that they can arbitrate both the reference-through-autograd
(`oracle/ref_exec.py`, `oracle/port_torch.py`) and the CUDA kernels.

Pinned by `tests/test_oracle.py` against `tests/golden/*.npz` (outputs of the reference's own
source executed through `oracle/tf_shim`; generator: `oracle/make_golden.py`).
"""
import numpy as np

F = np.float64


def _f(a):
    return np.asarray(a, dtype=F)


# --------------------------------------------------------------------------------------------
# cost matrices
# --------------------------------------------------------------------------------------------
def cost_xy(x, y, scaling_coef):
    """gan_utils.py:6-18 — C_ij = s * sum_{t,d} (x_itd - y_jtd)^2, direct (difference) form."""
    x = _f(x).reshape(len(x), -1)
    y = _f(y).reshape(len(y), -1)
    C = np.empty((x.shape[0], y.shape[0]), dtype=F)
    for i in range(x.shape[0]):          # row at a time: never materialises [B,B,T,D]
        d = x[i][None, :] - y
        C[i] = np.einsum("jk,jk->j", d, d)
    return C * scaling_coef


def martingale_operands(h, M):
    """gan_utils.py:34-35 — ht = h[:, :-1] and DeltaM = M[:, 1:] - M[:, :-1], flattened to [B,(T-1)J]."""
    h = _f(h)
    M = _f(M)
    Ht = h[:, :-1, :].reshape(h.shape[0], -1)
    dM = (M[:, 1:, :] - M[:, :-1, :]).reshape(M.shape[0], -1)
    return Ht, dM


def martingale_cost(h, M, scaling_coef):
    """gan_utils.py:37-38 — C_hM[i,j] = s * sum_{t<T-1,k} h[i,t,k] * DeltaM[j,t,k]
    (h indexed by the ROW sample, DeltaM by the COLUMN sample: SURVEY.md §0.5)."""
    Ht, dM = martingale_operands(h, M)
    return scaling_coef * (Ht @ dM.T)


def modified_cost(x, y, h, M, scaling_coef):
    """gan_utils.py:21-43."""
    return cost_xy(x, y, scaling_coef) + martingale_cost(h, M, scaling_coef)


def bi_causal_modified_cost(x, y, hy, Mx, hx, My, scaling_coef):
    """gan_utils.py:46-72 — modified_cost plus a second s * hx[i] . DeltaMy[j] term."""
    return (cost_xy(x, y, scaling_coef) + martingale_cost(hy, Mx, scaling_coef)
            + martingale_cost(hx, My, scaling_coef))


# --------------------------------------------------------------------------------------------
# log-domain Sinkhorn
# --------------------------------------------------------------------------------------------
def _lse(A, axis):
    m = A.max(axis=axis, keepdims=True)
    return (m + np.log(np.exp(A - m).sum(axis=axis, keepdims=True))).squeeze(axis)


def sinkhorn_forward(C, epsilon=1.0, L=100, Lmin=100, thresh=1e-2, exit_on_index=False):
    """gan_utils.py:138-165 (compute_sinkhorn) and :86-121 (benchmark_sinkhorn).

    u^k_i = a - eps*LSE_j((v^{k-1}_j - C_ij)/eps),  v^k_j = a - eps*LSE_i((u^k_i - C_ij)/eps),
    a = eps*log(1/B)  (the reference's `eps*(log mu - LSE((-C+u+v^T)/eps)) + u` — the old potential
    cancels).  Early exit: after an iteration, if thresh > sum|u - u_prev| and
    (#iterations done >= Lmin) [compute_sinkhorn, :157-160]  or  (0-based index >= Lmin)
    [benchmark_sinkhorn, :114-117; `exit_on_index=True`].
    Returns (cost, u_hist [nits+1,B], v_hist [nits+1,B], nits); hist[0] = 0.
    """
    C = _f(C)
    B = C.shape[0]
    a = epsilon * np.log(1.0 / B)
    u = np.zeros(B, dtype=F)
    v = np.zeros(B, dtype=F)
    uh, vh = [u], [v]
    nits = 0
    for i in range(int(L)):
        u1 = u
        u = a - epsilon * _lse((v[None, :] - C) / epsilon, axis=1)
        v = a - epsilon * _lse((u[:, None] - C) / epsilon, axis=0)
        uh.append(u)
        vh.append(v)
        nits += 1
        err = np.abs(u - u1).sum()
        if thresh > err and ((i >= Lmin) if exit_on_index else (nits >= Lmin)):
            break
    pi = np.exp((u[:, None] + v[None, :] - C) / epsilon)
    cost = float((pi * C).sum())
    return cost, np.stack(uh), np.stack(vh), nits


def sinkhorn_backward(C, epsilon, u_hist, v_hist, nits, gbar=1.0):
    """Reverse mode through the `nits` executed iterations of gan_utils.py:151-164 (the reference
    has no stop_gradient; TF's tape differentiates the unrolled loop and the sharp cost sum(pi*C)).
    Returns dcost/dC * gbar.  SURVEY.md Appendix A."""
    C = _f(C)
    B = C.shape[0]
    a = epsilon * np.log(1.0 / B)
    u, v = u_hist[nits], v_hist[nits]
    pi = np.exp((u[:, None] + v[None, :] - C) / epsilon)
    Cbar = pi * (1.0 - C / epsilon)
    ub = (pi * C).sum(axis=1) / epsilon
    vb = (pi * C).sum(axis=0) / epsilon
    for k in range(nits, 0, -1):
        Pv = np.exp((u_hist[k][:, None] + v_hist[k][None, :] - a - C) / epsilon)
        Cbar += Pv * vb[None, :]
        ub = ub - Pv @ vb
        Pu = np.exp((u_hist[k][:, None] + v_hist[k - 1][None, :] - a - C) / epsilon)
        Cbar += Pu * ub[:, None]
        vb = -(Pu.T @ ub)
        ub = np.zeros(B, dtype=F)
    return Cbar * gbar


def cost_backward(Cbar, x, y, scaling_coef, same=False):
    """Adjoint of cost_xy (gan_utils.py:14-17): gx_i = 2s * sum_j Cbar_ij (x_i - y_j),
    gy_j = 2s * sum_i Cbar_ij (y_j - x_i).  `same=True` (x is y) returns the sum of both roles."""
    xs, ys = np.shape(x), np.shape(y)
    X = _f(x).reshape(xs[0], -1)
    Y = _f(y).reshape(ys[0], -1)
    gx = 2.0 * scaling_coef * (Cbar.sum(axis=1)[:, None] * X - Cbar @ Y)
    gy = 2.0 * scaling_coef * (Cbar.sum(axis=0)[:, None] * Y - Cbar.T @ X)
    if same:
        return (gx + gy).reshape(xs)
    return gx.reshape(xs), gy.reshape(ys)


def martingale_backward(Cbar, h, M, scaling_coef):
    """Adjoint of martingale_cost: gh[:, :-1] = s * Cbar @ DeltaM, gh[:, -1] = 0;
    gDelta = s * Cbar^T @ ht;  gM[:, 1:] += gDelta, gM[:, :-1] -= gDelta."""
    h = _f(h)
    M = _f(M)
    Bh, T, J = h.shape
    Ht, dM = martingale_operands(h, M)
    gh = np.zeros_like(h)
    gh[:, :-1, :] = (scaling_coef * (Cbar @ dM)).reshape(Bh, T - 1, J)
    gD = (scaling_coef * (Cbar.T @ Ht)).reshape(M.shape[0], T - 1, J)
    gM = np.zeros_like(M)
    gM[:, 1:, :] += gD
    gM[:, :-1, :] -= gD
    return gh, gM


# --------------------------------------------------------------------------------------------
# entry points with gradients
# --------------------------------------------------------------------------------------------
def compute_sinkhorn(x, y, hy, Mx, scaling_coef, hx=None, My=None, epsilon=1.0, L=100,
                     bi_causal=False, grad=False):
    """gan_utils.py:124-165.  With grad=True also returns d/d{x, y, hy, Mx[, hx, My]}."""
    if bi_causal:
        C = bi_causal_modified_cost(x, y, hy, Mx, hx, My, scaling_coef)
    else:
        C = modified_cost(x, y, hy, Mx, scaling_coef)
    cost, uh, vh, nits = sinkhorn_forward(C, epsilon, L, Lmin=100, thresh=1e-2)
    if not grad:
        return cost
    Cbar = sinkhorn_backward(C, epsilon, uh, vh, nits)
    gx, gy = cost_backward(Cbar, x, y, scaling_coef)
    ghy, gMx = martingale_backward(Cbar, hy, Mx, scaling_coef)
    out = {"x": gx, "y": gy, "hy": ghy, "Mx": gMx, "C": C, "Cbar": Cbar,
           "u": uh[nits], "v": vh[nits], "nits": nits}
    if bi_causal:
        out["hx"], out["My"] = martingale_backward(Cbar, hx, My, scaling_coef)
    return cost, out


def benchmark_sinkhorn(x, y, scaling_coef, epsilon=1.0, L=10, Lmin=10):
    """gan_utils.py:75-121 — plain cost_xy Sinkhorn; the break tests the 0-based index."""
    C = cost_xy(x, y, scaling_coef)
    return sinkhorn_forward(C, epsilon, L, Lmin=Lmin, thresh=1e-2, exit_on_index=True)[0]


def compute_sinkhorn_loss(f_real, f_fake, scaling_coef, sinkhorn_eps, sinkhorn_l, h_fake, m_real,
                          h_real, m_fake, video=True, grad=False):
    """gan_utils.py:204-227.  `sinkhorn_eps`/`sinkhorn_l` land in compute_sinkhorn's `hx`/`My`
    slots (:221-223 vs :124) and are ignored: the solve always runs eps=1.0, L=100.
    The `video=True` transpose (:217-220) only permutes summed axes, so it is skipped."""
    del sinkhorn_eps, sinkhorn_l, video
    terms = {}
    g = {}
    for name, (x, y, h, M) in {"xy": (f_real, f_fake, h_fake, m_real),
                               "xx": (f_real, f_real, h_real, m_real),
                               "yy": (f_fake, f_fake, h_fake, m_fake)}.items():
        if grad:
            terms[name], g[name] = compute_sinkhorn(x, y, h, M, scaling_coef, grad=True)
        else:
            terms[name] = compute_sinkhorn(x, y, h, M, scaling_coef)
    loss = 2.0 * terms["xy"] - terms["xx"] - terms["yy"]
    if not grad:
        return loss
    grads = {
        "f_real": 2.0 * g["xy"]["x"] - (g["xx"]["x"] + g["xx"]["y"]),
        "f_fake": 2.0 * g["xy"]["y"] - (g["yy"]["x"] + g["yy"]["y"]),
        "h_fake": 2.0 * g["xy"]["hy"] - g["yy"]["hy"],
        "m_real": 2.0 * g["xy"]["Mx"] - g["xx"]["Mx"],
        "h_real": -g["xx"]["hy"],
        "m_fake": -g["yy"]["Mx"],
    }
    detail = {"loss_xy": terms["xy"], "loss_xx": terms["xx"], "loss_yy": terms["yy"]}
    for name in ("xy", "xx", "yy"):
        for k in ("C", "Cbar", "u", "v"):
            detail[f"{k}_{name}"] = g[name][k]
    return loss, grads, detail


def compute_N(M):
    """gan_utils.py:168-176."""
    M = _f(M)
    return M[:, 1:] - M[:, :-1]


def martingale_regularization(M, reg_lam, scaling_coef, grad=False):
    """gan_utils.py:179-201 — p_M = lam * s * sum_{t,j} | (1/m) sum_i DeltaM_itj / (std_j + 1e-6) |,
    std_j the POPULATION std of M[:, :, j] over (batch, time).  grad: d p_M / d M."""
    M = _f(M)
    m, T, J = M.shape
    N = M[:, 1:, :] - M[:, :-1, :]
    mean = M.mean(axis=(0, 1))
    sig = np.sqrt(((M - mean) ** 2).mean(axis=(0, 1)))
    den = sig + 1e-6
    A = N.sum(axis=0) / m / den                      # [T-1, J]
    pm = reg_lam * scaling_coef * np.abs(A).sum()
    if not grad:
        return float(pm)
    sgn = np.sign(A)
    w = reg_lam * scaling_coef
    gN = np.broadcast_to(w * sgn / (m * den), N.shape)   # via the numerator
    gM = np.zeros_like(M)
    gM[:, 1:, :] += gN
    gM[:, :-1, :] -= gN
    # via sigma_j:  dA_tj/dsig_j = -A_tj/den_j ;  dsig_j/dM_itj = (M_itj - mean_j)/(m*T*sig_j)
    gsig = w * (sgn * (-A / den)).sum(axis=0)         # [J]
    with np.errstate(divide="ignore", invalid="ignore"):
        gM += np.where(sig > 0, gsig / (m * T * sig), 0.0) * (M - mean)
    return float(pm), gM


# --------------------------------------------------------------------------------------------
# Gaussian kernel smoothing (data_utils.py:478-586)
# --------------------------------------------------------------------------------------------
def gaussian_kernel1d(radius, sigma):
    """data_utils.py:483-491 — w_k = exp(-k^2 / (2 sigma^2)) / sum, k in [-r, r]."""
    k = np.arange(-radius, radius + 1, dtype=F)
    w = np.exp(-0.5 / (sigma * sigma) * k ** 2)
    return w / w.sum()


def gaussian_kernel3d(radius, sigma):
    """data_utils.py:493-501 — equals w (x) w (x) w."""
    w = gaussian_kernel1d(radius, sigma)
    return np.einsum("a,b,c->abc", w, w, w)[:, :, :, None, None]


def _reflect_index(n, r):
    idx = np.arange(-r, n + r)
    idx = np.where(idx < 0, -idx, idx)
    return np.where(idx >= n, 2 * (n - 1) - idx, idx)


def _filter_matrix(n, radius, sigma):
    """[n,n] matrix A with (A z)_p = sum_k w_k z_reflect(p+k): REFLECT pad (data_utils.py:513)
    followed by a VALID cross-correlation (:515)."""
    if n <= radius:
        raise ValueError("REFLECT padding needs dim > radius")
    w = gaussian_kernel1d(radius, sigma)
    idx = _reflect_index(n, radius)
    A = np.zeros((n, n), dtype=F)
    for p in range(n):
        for k in range(2 * radius + 1):
            A[p, idx[p + k]] += w[k]
    return A


def _apply_axis(x, A, axis):
    return np.moveaxis(np.tensordot(A, np.moveaxis(x, axis, 0), axes=(1, 0)), 0, axis)


def _smooth(x, mats, grad_out=None):
    """out = conv(x) / max(conv(x)); optional VJP with TF's reduce_max gradient (split equally
    among tied arg-max elements)."""
    x = _f(x)
    z = x
    for axis, A in mats:
        z = _apply_axis(z, A, axis)
    m = z.max()
    out = z / m
    if grad_out is None:
        return out
    g = _f(grad_out)
    gz = g / m
    gm = -(g * z).sum() / (m * m)
    tie = (z == m)
    gz = gz + gm * tie / tie.sum()
    gx = gz
    for axis, A in reversed(mats):
        gx = _apply_axis(gx, A.T, axis)
    return out, gx


def temporal_convolution(x, sigma, radius=3, grad_out=None):
    """data_utils.py:503-521 — x [B,H,T,W,C]: 7-tap REFLECT filter along T, divided by the global
    max of the filtered tensor."""
    T = np.shape(x)[2]
    return _smooth(x, [(2, _filter_matrix(T, radius, sigma))], grad_out)


def gaussian_convolution3D(x, sigma, radius=3, grad_out=None):
    """data_utils.py:552-582 — per channel REFLECT-pad (T,H,W) by r, 7^3 VALID conv3d with
    w(x)w(x)w, divide by global max.  Separable: one 7-tap pass per axis H(1), T(2), W(3)."""
    _, H, T, W, _ = np.shape(x)
    mats = [(1, _filter_matrix(H, radius, sigma)), (2, _filter_matrix(T, radius, sigma)),
            (3, _filter_matrix(W, radius, sigma))]
    return _smooth(x, mats, grad_out)


def annealing_sigma(init_sigma, step, decay_steps=500, decay_rate=0.975):
    """data_utils.py:584-586."""
    return init_sigma * decay_rate ** (step / decay_steps)

"""STUDY CODE (not product, not oracle): a restatement of M.J.D. Powell's BOBYQA (report DAMTP 2009/NA06, "The BOBYQA algorithm
for bound constrained optimization without derivatives") for SMALL n, used only to bound the distance between the repository's
Nelder-Mead stand-in and the class of method the reference calls (nlopt 2.6.1 LN_BOBYQA, source/pmvs/optim.cpp:621-644).

What is Powell's, restated: npt = 2n+1 interpolation points; the quadratic model is updated by the least-Frobenius-norm change of
its Hessian subject to the interpolation conditions (the KKT system [[A, X^T], [X, 0]] with A_ij = (y_i . y_j)^2 / 2); the point
that leaves the set maximises |sigma_k| * max(1, (|y_k - x_opt| / Delta)^4) with sigma = alpha beta + tau^2; trust-region steps
inside the box; an alternative "geometry" step that maximises the Lagrange function of a far point over lines through x_opt and
the other points plus a projected-gradient step (ALTMOV); the rho ladder (rho <- 0.1 rho, or sqrt(rho rho_end), or rho_end)
with Delta <- max(rho_old / 2, rho_new); the 0.1 / 0.7 ratio thresholds for Delta.
What is simplified (n = 3 makes it affordable): the inverse KKT matrix is recomputed densely every iteration instead of being
updated through BMAT / ZMAT; the trust-region subproblem is solved by an eigen-decomposition with an active-set loop on the
bounds instead of TRSBOX's truncated conjugate gradients; there is no origin shift and no RESCUE (double precision and n = 3 do
not need them).  nlopt's wrapper is followed for the scaling: variables are divided by s_i = dx_i / dx_0 (dx = nlopt's default
initial steps), rho_beg = dx_0, rho_end = xtol_rel * rho_beg."""
import numpy as np


def nlopt_default_steps(x, lb, ub):
    """nlopt_set_default_initial_step (nlopt 2.6.1 src/api/options.c)"""
    dx = np.zeros(len(x))
    for i in range(len(x)):
        step = np.inf
        if np.isfinite(ub[i]) and np.isfinite(lb[i]) and (ub[i] - lb[i]) * 0.25 < step and ub[i] > lb[i]:
            step = (ub[i] - lb[i]) * 0.25
        if np.isfinite(ub[i]) and ub[i] - x[i] < step and ub[i] > x[i]:
            step = (ub[i] - x[i]) * 0.75
        if np.isfinite(lb[i]) and x[i] - lb[i] < step and x[i] > lb[i]:
            step = (x[i] - lb[i]) * 0.75
        if not np.isfinite(step):
            if np.isfinite(ub[i]) and abs(ub[i] - x[i]) < abs(step):
                step = (ub[i] - x[i]) * 1.1
            if np.isfinite(lb[i]) and abs(x[i] - lb[i]) < abs(step):
                step = (x[i] - lb[i]) * 1.1
        if not np.isfinite(step) or abs(step) < 1e-300:
            step = x[i]
        if not np.isfinite(step) or step == 0.0:
            step = 1.0
        dx[i] = step
    return dx


class _Model:
    """quadratic c + g.x + x^T H x / 2 around the origin of the scaled, shifted variables"""

    def __init__(self, n):
        self.c = 0.0
        self.g = np.zeros(n)
        self.H = np.zeros((n, n))

    def __call__(self, x):
        return self.c + self.g @ x + 0.5 * x @ self.H @ x

    def grad(self, x):
        return self.g + self.H @ x


def _kkt_inverse(Y):
    npt, n = Y.shape
    A = 0.5 * (Y @ Y.T) ** 2
    X = np.vstack([np.ones(npt), Y.T])           # (n+1, npt)
    W = np.zeros((npt + n + 1, npt + n + 1))
    W[:npt, :npt] = A
    W[:npt, npt:] = X.T
    W[npt:, :npt] = X
    return np.linalg.pinv(W)


def _min_frobenius_update(model, Hinv, Y, resid):
    """add to `model` the quadratic D with D(y_j) = resid_j and the smallest ||grad^2 D||_F"""
    npt, n = Y.shape
    sol = Hinv @ np.concatenate([resid, np.zeros(n + 1)])
    lam, c, g = sol[:npt], sol[npt], sol[npt + 1:]
    model.c += c
    model.g += g
    model.H += (Y.T * lam) @ Y


def _lagrange(Hinv, Y, k):
    npt, n = Y.shape
    col = Hinv[:, k]
    m = _Model(n)
    m.c = col[npt]
    m.g = col[npt + 1:].copy()
    m.H = (Y.T * col[:npt]) @ Y
    return m


def _trust_step(g, H, delta, lo, hi):
    """min g.d + d^T H d / 2  s.t. |d| <= delta, lo <= d <= hi (lo <= 0 <= hi).  Eigen-decomposition in the free subspace, variables
    that hit a bound are fixed there and the rest re-solved (at most n passes)."""
    n = len(g)
    free = np.ones(n, bool)
    d = np.zeros(n)
    for _ in range(n + 1):
        idx = np.where(free)[0]
        if len(idx) == 0:
            break
        fixed = d.copy(); fixed[free] = 0.0
        gf = (g + H @ fixed)[idx]
        Hf = H[np.ix_(idx, idx)]
        rad2 = delta * delta - fixed @ fixed
        if rad2 <= 0:
            break
        rad = np.sqrt(rad2)
        w, V = np.linalg.eigh(Hf)
        gt = V.T @ gf

        def step(mu):
            return -V @ (gt / (w + mu))
        if w[0] > 1e-14 and np.linalg.norm(step(0.0)) <= rad:
            df = step(0.0)
        else:
            lo_mu = max(0.0, -w[0]) + 1e-14
            hi_mu = lo_mu + np.linalg.norm(gf) / max(rad, 1e-300) + abs(w).max() + 1.0
            if np.linalg.norm(gf) < 1e-300:
                df = V[:, 0] * rad if w[0] < 0 else np.zeros(len(idx))
            else:
                for _it in range(100):
                    mu = 0.5 * (lo_mu + hi_mu)
                    if np.linalg.norm(step(mu)) > rad:
                        lo_mu = mu
                    else:
                        hi_mu = mu
                df = step(hi_mu)
        trial = fixed.copy(); trial[idx] = df
        viol = (trial < lo - 1e-15) | (trial > hi + 1e-15)
        if not viol.any():
            d = trial
            break
        # the most violated variable goes to its bound and stays there
        excess = np.maximum(lo - trial, trial - hi)
        j = int(np.argmax(excess))
        d = fixed
        d[j] = lo[j] if trial[j] < lo[j] else hi[j]
        free[j] = False
        if not free.any():
            break
    return np.clip(d, lo, hi)


def bobyqa(f, x0, lb, ub, xtol_rel=1e-7, maxeval=1000, dx=None):
    """-> (x, fx, evaluations, status) ; status 'xtol' or 'maxeval'"""
    x0 = np.asarray(x0, float); lb = np.asarray(lb, float); ub = np.asarray(ub, float)
    n = len(x0)
    npt = 2 * n + 1
    dx = nlopt_default_steps(x0, lb, ub) if dx is None else np.asarray(dx, float)
    s = np.ones(n)
    if not np.all(dx == dx[0]):
        s[1:] = np.abs(dx[1:] / dx[0])           # nlopt_compute_rescaling
    sl, su = lb / s, ub / s
    rhobeg = abs(dx[0] / s[0])
    for j in range(n):
        if np.isfinite(su[j] - sl[j]):
            rhobeg = min(rhobeg, 0.5 * (su[j] - sl[j]))
    rhoend = xtol_rel * rhobeg
    base = np.clip(x0 / s, sl, su)
    # BOBYQA moves the start so that it is either on a bound or at least rhobeg away from it
    for j in range(n):
        if base[j] - sl[j] < rhobeg and base[j] > sl[j]:
            base[j] = sl[j] + rhobeg if sl[j] + rhobeg <= su[j] else base[j]
        if su[j] - base[j] < rhobeg and base[j] < su[j]:
            base[j] = su[j] - rhobeg if su[j] - rhobeg >= sl[j] else base[j]
    evals = [0]

    def F(y):                                     # y = offset from base in scaled variables
        evals[0] += 1
        return float(f((base + y) * s))
    lo, hi = sl - base, su - base                 # bounds on the offsets
    # PRELIM: x0, x0 +- rhobeg e_i (a step that would leave the box is replaced by 2 rhobeg the other way)
    Y = np.zeros((npt, n))
    for i in range(n):
        a, b = rhobeg, -rhobeg
        if hi[i] < a:
            a, b = -rhobeg, -2 * rhobeg
        elif lo[i] > b:
            a, b = rhobeg, 2 * rhobeg
        Y[1 + i, i] = a
        Y[1 + n + i, i] = min(max(b, lo[i]), hi[i])
    fv = np.array([F(y) for y in Y])
    model = _Model(n)
    Hinv = _kkt_inverse(Y)
    _min_frobenius_update(model, Hinv, Y, fv.copy())
    kopt = int(np.argmin(fv))
    rho, delta = rhobeg, rhobeg
    status = "xtol"
    while True:
        if evals[0] >= maxeval:
            status = "maxeval"
            break
        xopt = Y[kopt].copy()
        gopt = model.grad(xopt)
        d = _trust_step(gopt, model.H, delta, lo - xopt, hi - xopt)
        dnorm = np.linalg.norm(d)
        dist2 = ((Y - xopt) ** 2).sum(1)
        reduce_rho = False
        geometry = None
        ratio = -1.0
        if dnorm < 0.5 * rho:
            far = max((2 * delta) ** 2, (10 * rho) ** 2)
            k = int(np.argmax(dist2))
            if dist2[k] > far:
                geometry = k
            else:
                reduce_rho = True
        else:
            xnew = np.clip(xopt + d, lo, hi)
            pred = model(xopt) - model(xnew)
            fnew = F(xnew)
            ratio = (fv[kopt] - fnew) / pred if pred > 0 else -1.0
            if ratio <= 0.1:
                delta = min(0.5 * delta, dnorm)
            elif ratio <= 0.7:
                delta = max(0.5 * delta, dnorm)
            else:
                delta = max(0.5 * delta, 2 * dnorm)
            if delta <= 1.5 * rho:
                delta = rho
            # which point leaves: sigma_k = alpha_k beta + tau_k^2 weighted by the distance from the (new) best point
            w = np.concatenate([0.5 * (Y @ xnew) ** 2, [1.0], xnew])
            Hw = Hinv @ w
            tau = Hw[:npt]
            beta = 0.5 * (xnew @ xnew) ** 2 - w @ Hw
            alpha = np.diag(Hinv)[:npt]
            sigma = alpha * beta + tau ** 2
            centre = xnew if fnew < fv[kopt] else xopt
            wdist = np.maximum(1.0, (((Y - centre) ** 2).sum(1) / max(delta * delta, 1e-300)) ** 2)
            score = np.abs(sigma) * wdist
            if not fnew < fv[kopt]:
                score[kopt] = -1.0
            knew = int(np.argmax(score))
            resid = np.zeros(npt)
            Y[knew] = xnew
            fv[knew] = fnew
            Hinv = _kkt_inverse(Y)
            resid[knew] = fnew - model(xnew)
            _min_frobenius_update(model, Hinv, Y, resid)
            if fnew < fv[kopt] or knew == kopt:
                kopt = int(np.argmin(fv))
            if ratio >= 0.1:
                continue
            xopt = Y[kopt]
            dist2 = ((Y - xopt) ** 2).sum(1)
            far = max((2 * delta) ** 2, (10 * rho) ** 2)
            k = int(np.argmax(dist2))
            if dist2[k] > far:
                geometry = k
            elif ratio > 0 or max(delta, dnorm) > rho:
                continue
            else:
                reduce_rho = True
        if geometry is not None:
            if evals[0] >= maxeval:
                status = "maxeval"
                break
            k = geometry
            xopt = Y[kopt].copy()
            adelt = max(min(0.1 * np.sqrt(dist2[k]), delta), rho)
            lag = _lagrange(Hinv, Y, k)
            cands = []
            for j in range(npt):                  # lines through x_opt and the other points, both ways
                if j == kopt:
                    continue
                v = Y[j] - xopt
                nv = np.linalg.norm(v)
                if nv == 0:
                    continue
                for sgn in (1.0, -1.0):
                    cands.append(np.clip(xopt + sgn * adelt * v / nv, lo, hi))
            gl = lag.grad(xopt)                   # projected-gradient (Cauchy) steps of the Lagrange function
            for sgn in (1.0, -1.0):
                v = sgn * gl.copy()
                v[(xopt <= lo) & (v < 0)] = 0.0
                v[(xopt >= hi) & (v > 0)] = 0.0
                nv = np.linalg.norm(v)
                if nv > 0:
                    cands.append(np.clip(xopt + adelt * v / nv, lo, hi))
            vals = [abs(lag(c)) for c in cands]
            xnew = cands[int(np.argmax(vals))]
            if np.linalg.norm(xnew - xopt) < 1e-3 * adelt:
                reduce_rho = True
            else:
                fnew = F(xnew)
                resid = np.zeros(npt)
                resid[k] = fnew - model(xnew)
                Y[k] = xnew
                fv[k] = fnew
                Hinv = _kkt_inverse(Y)
                _min_frobenius_update(model, Hinv, Y, resid)
                kopt = int(np.argmin(fv))
                continue
        if reduce_rho:
            if rho <= rhoend:
                break
            old = rho
            r = rho / rhoend
            rho = rhoend if r <= 16 else (np.sqrt(r) * rhoend if r <= 250 else 0.1 * rho)
            delta = max(0.5 * old, rho)
    k = int(np.argmin(fv))
    return (base + Y[k]) * s, float(fv[k]), evals[0], status


if __name__ == "__main__":                       # sanity: smooth test functions
    rng = np.random.default_rng(0)
    rosen = lambda x: 100 * (x[1] - x[0] ** 2) ** 2 + (1 - x[0]) ** 2 + (x[2] - 0.5) ** 2
    x, fx, ne, st = bobyqa(rosen, [-1.2, 1.0, 0.0], [-np.inf] * 3, [np.inf] * 3, xtol_rel=1e-8)
    print("rosenbrock+1:", x, fx, ne, st)
    Q = rng.normal(size=(3, 3)); Q = Q @ Q.T + 0.1 * np.eye(3)
    quad = lambda x: 0.5 * (x - 1) @ Q @ (x - 1)
    x, fx, ne, st = bobyqa(quad, [0.0, 0.0, 0.0], [-np.inf, -0.5, -0.5], [np.inf, 0.5, 0.5], xtol_rel=1e-8)
    print("bounded quadratic:", x, fx, ne, st)

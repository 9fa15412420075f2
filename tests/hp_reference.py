"""A 50-digit (mpmath) evaluation of the exact-GP posterior of one arm — SingleTaskGP with Normalize + Standardize,
ScaleKernel(Matern-5/2 ARD), constant mean, homoskedastic noise — written independently of
oracle/gp_oracle.py: dense inverse by LU instead of Cholesky, the correlation from the general Bessel-K Matern
formula 2^(1-nu)/Gamma(nu) (sqrt(2 nu) r)^nu K_nu(sqrt(2 nu) r) instead of its nu = 5/2 closed form.  Test
infrastructure: it pins the NUMERICS of the oracle (and of the CUDA kernels) on an ill-conditioned problem."""
import mpmath as mp
import torch


def ill_conditioned_problem():
    g = torch.Generator().manual_seed(12)
    n, d, q = 9, 2, 5
    X = torch.rand(n, d, dtype=torch.float64, generator=g)
    X[1] = X[0] + 1e-3                      # two nearly coincident inputs: cond(K^) > 1e4
    Y = torch.sin(6 * X).sum(-1, keepdim=True) + 0.05 * torch.randn(n, 1, dtype=torch.float64, generator=g)
    ls = torch.tensor([0.35, 0.8], dtype=torch.float64)
    Xq = torch.rand(q, d, dtype=torch.float64, generator=g)
    Xq[0] = X[3]                            # a query on top of a training input (r = 0)
    return X, Y, ls, 1.7, 2e-4, 0.15, Xq


def mp_posterior(X, Y, ls, osc, noise, mconst, Xq, digits=50):
    """Returns (mean [q], variance [q]) as float64 tensors rounded from `digits`-digit arithmetic, cond_1(K^) and the
    prior variance scale s^2 o."""
    mp.mp.dps = digits
    n, d = X.shape
    q = Xq.shape[0]
    rng = X.max(0).values - X.min(0).values

    def M(t):
        return mp.matrix([[mp.mpf(float(v)) for v in row] for row in t.tolist()])

    Xm, Qm, Ym = M(X), M(Xq), M(Y)
    nu = mp.mpf(5) / 2

    def corr(a, b):
        r = mp.sqrt(sum(((a[k] - b[k]) / (mp.mpf(float(rng[k])) * mp.mpf(float(ls[k])))) ** 2 for k in range(d)))
        if r == 0:
            return mp.mpf(1)
        z = mp.sqrt(2 * nu) * r
        return 2 ** (1 - nu) / mp.gamma(nu) * z ** nu * mp.besselk(nu, z)

    def row(A, i):
        return [A[i, k] for k in range(d)]

    ybar = sum(Ym[i, 0] for i in range(n)) / n
    ysd = mp.sqrt(sum((Ym[i, 0] - ybar) ** 2 for i in range(n)) / (n - 1))
    K = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            K[i, j] = mp.mpf(osc) * corr(row(Xm, i), row(Xm, j)) + (mp.mpf(noise) if i == j else 0)
    Kinv = mp.inverse(K)
    resid = mp.matrix([(Ym[i, 0] - ybar) / ysd - mp.mpf(mconst) for i in range(n)])
    alpha = Kinv * resid
    cond = float(mp.norm(K, 1) * mp.norm(Kinv, 1))
    mean, var = [], []
    for c in range(q):
        ks = mp.matrix([mp.mpf(osc) * corr(row(Qm, c), row(Xm, j)) for j in range(n)])
        mu = mp.mpf(mconst) + sum(ks[j] * alpha[j] for j in range(n))
        v = mp.mpf(osc) - sum(ks[i] * sum(Kinv[i, j] * ks[j] for j in range(n)) for i in range(n))
        mean.append(float(ybar + ysd * mu))
        var.append(float(ysd * ysd * v))
    return (torch.tensor(mean, dtype=torch.float64), torch.tensor(var, dtype=torch.float64), cond,
            float(ysd * ysd * mp.mpf(osc)))

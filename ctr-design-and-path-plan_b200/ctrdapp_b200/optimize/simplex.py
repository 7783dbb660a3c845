"""Nelder-Mead with concurrent evaluation of the points an iteration may need.

The reference hands `scipy.optimize.minimize(method='Nelder-Mead')` an objective that is a WHOLE RRT run per
design q (neldermead.py:83-135) and SciPy evaluates one point at a time.  The simplex method has little, but
real, parallelism:
  * the dim + 1 vertices of the initial simplex are independent (6 full RRT runs for a 5-parameter design);
  * within an iteration the reflected, expanded and both contracted points are all functions of the current
    simplex alone, so they can be evaluated speculatively side by side;
  * a shrink step evaluates dim new vertices at once.
`nelder_mead` takes `eval_many(list of x) -> list of f` -- one vertex per GPU in `DistributedEvaluator` -- and
otherwise follows SciPy's algorithm (rho 1, chi 2, psi 0.5, sigma 0.5, its acceptance tests, its stable sort and
its termination rule max|x_i - x_0| <= xatol and max|f_0 - f_i| <= fatol), so that with a deterministic
objective the iterates, `nit` and the sequential evaluation count `nfev` equal SciPy's
(tests/test_optimize_host.py).  `nfev_speculative` counts what was evaluated on top of that.
"""
import numpy as np


class NelderMeadResult(dict):
    __getattr__ = dict.get


def nelder_mead(eval_many, initial_simplex, xatol=1e-4, fatol=1e-4, maxfev=None, maxiter=None, speculate=True, on_request=None):
    """speculate: False / 0 = SciPy's one point at a time; True / 3 = reflected + expanded + both contracted points side by
    side; 1 or 2 = that many of [expanded, contracted, inner-contracted] beside the reflected point (as many as there are
    idle ranks).  The result's `used_speculative` lists the speculative points the sequential algorithm then asked for.
    on_request(n_speculative) is called before every eval_many request with the number of its TRAILING points that are
    speculative (0 for the initial simplex, a shrink step or a point the algorithm asks for by itself)."""
    note = on_request or (lambda n_speculative: None)
    sim = np.array(initial_simplex, dtype=np.float64)
    nspec = 3 if speculate is True else int(speculate or 0)
    nspec = max(0, min(3, nspec))
    used = []
    n = sim.shape[1]
    if sim.shape[0] != n + 1:
        raise ValueError("`initial_simplex` should be an array of shape (N+1,N)")
    if maxfev is None and maxiter is None:
        maxfev = maxiter = n * 200
    maxfev = np.inf if maxfev is None else maxfev
    maxiter = np.inf if maxiter is None else maxiter
    rho, chi, psi, sigma = 1.0, 2.0, 0.5, 0.5
    nfev = spec = 0

    budget = int(min(n + 1, maxfev))
    fsim = np.full(n + 1, np.inf)
    note(0)
    fsim[:budget] = eval_many([sim[k] for k in range(budget)])
    nfev += budget
    order = np.argsort(fsim, kind="stable")
    sim, fsim = sim[order], fsim[order]
    nit = 1
    while nfev < maxfev and nit < maxiter:
        if np.max(np.abs(sim[1:] - sim[0])) <= xatol and np.max(np.abs(fsim[0] - fsim[1:])) <= fatol:
            break
        xbar = np.add.reduce(sim[:-1], 0) / n
        xr = (1 + rho) * xbar - rho * sim[-1]
        xe = (1 + rho * chi) * xbar - rho * chi * sim[-1]
        xc = (1 + psi * rho) * xbar - psi * rho * sim[-1]
        xcc = (1 - psi) * xbar + psi * sim[-1]
        cand = [("e", xe), ("c", xc), ("cc", xcc)][:nspec]
        note(len(cand))
        vals = eval_many([xr] + [x for _, x in cand])
        fxr = vals[0]
        lazy = {tag: (x, v) for (tag, x), v in zip(cand, vals[1:])}
        spec += len(cand)
        nfev += 1

        def need(tag, x):
            """value of a point the sequential algorithm evaluates now"""
            nonlocal nfev, spec
            nfev += 1
            if tag in lazy:
                spec -= 1
                used.append(np.array(lazy[tag][0]))
                return lazy[tag][1]
            note(0)
            return eval_many([x])[0]

        doshrink = False
        if fxr < fsim[0]:
            fxe = need("e", xe)
            if fxe < fxr:
                sim[-1], fsim[-1] = xe, fxe
            else:
                sim[-1], fsim[-1] = xr, fxr
        elif fxr < fsim[-2]:
            sim[-1], fsim[-1] = xr, fxr
        elif fxr < fsim[-1]:
            fxc = need("c", xc)
            if fxc <= fxr:
                sim[-1], fsim[-1] = xc, fxc
            else:
                doshrink = True
        else:
            fxcc = need("cc", xcc)
            if fxcc < fsim[-1]:
                sim[-1], fsim[-1] = xcc, fxcc
            else:
                doshrink = True
        if doshrink:
            for j in range(1, n + 1):
                sim[j] = sim[0] + sigma * (sim[j] - sim[0])
            note(0)
            fsim[1:] = eval_many([sim[j] for j in range(1, n + 1)])
            nfev += n
        order = np.argsort(fsim, kind="stable")
        sim, fsim = sim[order], fsim[order]
        nit += 1
    success = nfev < maxfev and nit < maxiter
    return NelderMeadResult(x=sim[0].copy(), fun=float(fsim[0]), nit=nit, nfev=nfev, nfev_speculative=spec, success=bool(success),
                            final_simplex=(sim, fsim), used_speculative=used)


class DistributedEvaluator:
    """eval_many over the ranks of a torch.distributed job (one process per GPU): point k of a request goes to
    rank k mod world, every rank evaluates its share with its own model / collision handles, and one all_gather
    of the float64 costs gives every rank the full answer, so all ranks take the same simplex decisions."""

    def __init__(self, func, device="cpu", group=None):
        self.func, self.device, self.group = func, device, group
        self.calls = 0

    def world_size(self):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return 1
        return dist.get_world_size(self.group)

    def rank(self):
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return 0
        return dist.get_rank(self.group)

    def owner_of_lowest(self, local_value):
        """rank holding the lowest of the ranks' local values (lowest rank on a tie); one all_gather of a float64"""
        import torch
        import torch.distributed as dist
        if self.world_size() == 1:
            return 0
        mine = torch.tensor([float(local_value)], dtype=torch.float64, device=self.device)
        out = torch.empty(self.world_size(), dtype=torch.float64, device=self.device)
        dist.all_gather_into_tensor(out, mine, group=self.group)
        return int(np.argmin(out.cpu().numpy()))

    def __call__(self, xs):
        import torch
        import torch.distributed as dist
        xs = [np.asarray(x, np.float64) for x in xs]
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(self.group) == 1:
            self.calls += len(xs)
            return [float(self.func(x)) for x in xs]
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        per = -(-len(xs) // world)
        mine = torch.full((per,), float("nan"), dtype=torch.float64, device=self.device)
        for slot in range(per):
            k = slot * world + rank
            if k < len(xs):
                mine[slot] = float(self.func(xs[k]))
                self.calls += 1
        out = torch.empty(world * per, dtype=torch.float64, device=self.device)
        dist.all_gather_into_tensor(out, mine, group=self.group)
        table = out.view(world, per).cpu().numpy()
        return [float(table[k % world, k // world]) for k in range(len(xs))]

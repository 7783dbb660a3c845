"""Design optimisation over q (mirror of ctrdapp/optimize/neldermead.py): the objective of a design is the best
cost of a whole RRT run with that design.  Differences from the reference, all in the driver, none in the
objective: the interactive `input()` prompts (neldermead.py:62-77) are replaced by the optional configuration key
`initial_simplex`; SciPy's one-point-at-a-time Nelder-Mead is replaced by `simplex.nelder_mead`, which evaluates
independent vertices concurrently -- one per GPU under torchrun."""
import time

import numpy as np

from ..model.model_factory import create_model
from ..model.strain_bases import max_from_base
from ..solve.solver_factory import create_solver
from .optimize_result import OptimizeResult
from .simplex import DistributedEvaluator, nelder_mead


class Optimizer:
    def __init__(self, heuristic_factory, collision_checker, initial_guess, configuration):
        self.tube_num = configuration.get("tube_number")
        self.precision = configuration.get("optimizer_precision")
        self.configuration = configuration
        self.heuristic_factory = heuristic_factory
        self.collision_checker = collision_checker
        self.initial_guess = initial_guess

    def find_min(self):
        raise NotImplementedError


def calculate_simplex(base_names, tube_radii, length, q_dof, initial_guess, e_max=0.05):
    """Initial simplex from the elastic strain limit (neldermead.py:143-191): every q is capped by what a strain
    of e_max allows for its base (2 e_max / diameter for a curvature), the first vertex is half of the caps, vertex
    i + 1 quarters cap i."""
    if initial_guess is None:
        initial_guess = np.ones(sum(q_dof)) * 100
    caps = []
    for i, r in enumerate(tube_radii):
        for k, val in enumerate(max_from_base(base_names[i], e_max / r, length[i], q_dof[i])):
            guess = initial_guess[sum(q_dof[0:i]) + k]
            caps.append(guess if val > guess else val)
    simplex = [[c / 2 for c in caps]]
    for i in range(len(caps)):
        vertex = list(caps)
        vertex[i] = vertex[i] / 4
        simplex.append(vertex)
    return simplex


class NelderMead(Optimizer):
    def __init__(self, heuristic_factory, collision_checker, initial_guess, configuration, device="cpu", group=None):
        super().__init__(heuristic_factory, collision_checker, initial_guess, configuration)
        self.solver_results = []
        self.best_solver = {'cost': 10e9, 'solver': None}
        self.count = 0
        self._eval = DistributedEvaluator(self.solver_heuristic, device=device, group=group)

    def find_min(self):
        conf = self.configuration
        radius = conf.get('tube_radius').get('outer')
        if radius[0] == 5:  # large radius for the single-tube FTL experiment (neldermead.py:36-37)
            radius = [0.9]
        init_simplex = conf.get('initial_simplex') or calculate_simplex(conf.get('strain_bases'), radius, conf.get('tube_lengths'),
                                                                        conf.get('q_dof'), self.initial_guess)
        start = time.time()
        # Speculative evaluation only pays when other ranks would idle: by default one speculative point per idle rank
        # (none in a single process, where the run is then SciPy's algorithm evaluation for evaluation, as the
        # reference's); 'speculate' in the configuration overrides (True = all three, or a count).
        world = self._eval.world_size()
        speculate = conf.get('speculate')
        if speculate is None:
            speculate = min(3, world - 1)
        res = nelder_mead(self._record, init_simplex, xatol=self.precision, fatol=self.precision,
                          maxfev=conf.get('optimize_iterations'), speculate=speculate, on_request=self._expect)
        # speculative points the algorithm then asked for are ordinary evaluations; the others stay tagged, so that
        # optimize_process.txt lists what the reference's sequential run would have listed
        used = {tuple(np.asarray(x, float)) for x in res.used_speculative}
        for entry in self.solver_results:
            if entry.get('speculative') and tuple(entry['q']) in used:
                entry['speculative'] = False
        # every rank ran its own share of the designs: the solver behind the overall best cost lives on ONE rank
        best_rank = self._eval.owner_of_lowest(self.best_solver.get('cost'))
        mine = self._eval.rank() == best_rank
        result = OptimizeResult(res.x, self.best_solver.get('solver') if mine else None, res.success, res.nfev, res.nit,
                                time.time() - start, [e for e in self.solver_results if not e.get('speculative')])
        result.num_speculative_eval = res.nfev_speculative
        result.speculative_results = [e for e in self.solver_results if e.get('speculative')]
        result.best_rank = best_rank          # the rank whose result.best_solver is set: call save_result there
        result.owns_best_solver = mine
        return result

    def _expect(self, n_speculative):
        self._n_speculative = n_speculative

    def _record(self, xs):
        costs = self._eval(xs)
        n_spec = getattr(self, "_n_speculative", 0)
        for k, (x, c) in enumerate(zip(xs, costs)):
            # the trailing n_spec points of the request are speculative until the algorithm asks for them (find_min
            # clears the tag then)
            self.solver_results.append({'q': list(np.asarray(x, float)), 'cost': c, 'speculative': k >= len(xs) - n_spec})
        return costs

    def solver_heuristic(self, x):
        """cost of one design: build the model, run the planner, take its best cost (neldermead.py:101-135)."""
        model = create_model(self.configuration, list(np.asarray(x, float)), ctx=getattr(self.collision_checker, "ctx", None))
        solver = create_solver(model, self.heuristic_factory, self.collision_checker, self.configuration)
        cost, _ = solver.get_best_cost()
        if cost < self.best_solver.get('cost'):
            self.best_solver = {'cost': cost, 'solver': solver}
        self.count += 1
        return cost

"""OptimizeResult (mirror of ctrdapp/optimize/optimize_result.py:9-103 without the matplotlib / pyvista plots):
writes tree.txt, solution_path.txt, optimize_result.txt and optimize_process.txt in the reference's formats."""


class OptimizeResult:
    def __init__(self, best_q, best_solver, success, num_function_eval, num_iterations, completion_time, optimize_process):
        self.best_q = best_q
        self.best_solver = best_solver
        self.success = success
        self.num_function_eval = num_function_eval
        self.num_iterations = num_iterations
        self.completion_time = completion_time
        self.optimize_process = optimize_process  # list of {q, cost} in evaluation order

    def save_result(self, output_dir):
        self.best_solver.save_tree(output_dir)
        self.best_solver.save_best_solution(output_dir)
        with open(output_dir / "optimize_result.txt", "w") as f:
            f.write(f"Optimizer Success: {self.success}\n")
            f.write(f"Goal Reached: {self.best_solver.found_solution}\n")
            f.write(f"Completion Time: {self.completion_time}\n")
            f.write(f"Number of Iterations: {self.num_iterations}\n")
            f.write(f"Number of Function Evals: {self.num_function_eval}\n")
            f.write("Optimized Q: " + ", ".join(str(v) for v in self.best_q) + "\n")
        with open(output_dir / "optimize_process.txt", "w") as f:
            for entry in self.optimize_process:
                f.write(", ".join(str(v) for v in entry.get("q")) + f" | {entry.get('cost')}\n")

"""Model definitions shared by the golden generator and the parity tests."""
MODELS = {
    "config_yaml": (["linear", "linear"], [2, 3], [0.02, 0.001, 0.05, 0.0001, 0.002], [120, 120], 1),
    "c3_three_tube": (["linear", "linear", "linear"], [2, 3, 3],
                      [0.02, 0.001, 0.05, 0.0001, 0.002, 0.03, 0.0002, 0.0005], [33, 66, 99], 1),
    "mixed_bases": (["quadratic", "pure_helix", "torsion_linear_helix"], [3, 2, 3],
                    [0.01, 1e-5, 2e-5, 0.03, 0.02, 0.05, 0.01, 3e-4], [20, 30.5, 45], 0.5),
    "full_base": (["full", "linear_helix"], [8, 2],
                  [0.01, 1e-4, 0.02, 1e-4, 1e-6, 0.015, 2e-4, 1e-6, 0.03, 5e-4], [40, 60], 2),
    "torsion_constant": (["torsion_helix", "constant", "full", "quadratic"], [3, 1, 5, 2],
                         [0.05, 0.02, 0.01, 0.0, 0.02, 0.01, 1e-4, 0.02, 2e-4, 0.03, 1e-5], [10, 20, 30, 40], 0.25),
}

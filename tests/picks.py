"""Scenario ids selected for rare exits of Optimize, shared by the CPU pins and the GPU parity tests."""

# scenarios of scenarios.generate(seed, 0, B, N) that leave Optimize through the gradient-norm test
# (ilqr_optimizer.cc:235-241) when the two cost tolerances are 0 -- found by scanning the oracle
GRAD_EXIT_PICKS = {(5, 30): [38, 404, 882, 996, 1307, 1487], (6, 50): [436, 632]}

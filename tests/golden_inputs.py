"""Seeded inputs of the golden vectors under tests/golden/: shared by oracle/gen_golden.py (which feeds them to the
reference's own kernels on a B200 and stores the outputs) and by the tests (which feed them to the oracle / the CUDA path)."""
import numpy as np

N_GRID = 16384      # samples of the hash-grid forward / SH golden
N_GRID_BWD = 2048   # samples of the hash-grid backward golden (first N_GRID_BWD of the same positions)


def grid_inputs(n_grid_params, seed=1234):
    rs = np.random.RandomState(seed)
    table = (rs.randn(n_grid_params) * 0.5).astype(np.float16)
    positions = rs.rand(N_GRID, 3).astype(np.float32)
    positions[0] = 0.0; positions[1] = 1.0; positions[2] = [0.5, 0.25, 0.75]; positions[3] = [1.0, 0.0, 1.0]
    dy = (rs.randn(32, N_GRID) * 0.01).astype(np.float16)   # [feature][sample], the reference's layout
    dirs = rs.rand(N_GRID, 3).astype(np.float32)
    return table, positions, dy, dirs

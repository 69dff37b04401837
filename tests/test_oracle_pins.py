"""Pins the oracle's F32 mul_mat to what the reference's own Test3 asserts (Test3/Program.cs:22-88): with F built by
the test's LCG and the sign pattern L-BFGS must converge to, F.x* reproduces the labels l; and the least-squares
minimiser of the test's objective, computed with the oracle's mul_mat, is within the test's 1e-2 of +-1."""
import numpy as np

from oracle import pyoracle as orc


def test3_problem():
    NP, NF = 1 << 12, 1 << 8
    nxt, r = 0, np.zeros(NP * NF, dtype=np.float32)
    for n in range(NP * NF):                       # xrand() with xsrand(0), Test3/Program.cs:98-107
        nxt = (nxt * 214013 + 2531011) & 0xFFFFFFFFFFFFFFFF
        r[n] = (nxt >> 16) & 0x7FFF
    l = np.where(np.arange(NP) < NP // 2, 1.0, -1.0).astype(np.float32)
    i = np.arange(NF)[None, :]
    ind = np.where(((l[:, None] > 0) & (i < NF // 2)) | ((l[:, None] < 0) & (i >= NF // 2)), 1.0, 0.0).astype(np.float32)
    noise = (r.reshape(NP, NF) / np.float32(32767) - np.float32(0.5)) * np.float32(0.1)
    return ((ind + noise) / np.float32(0.5 * NF)).astype(np.float32), l


def test_oracle_f32_mul_mat_reproduces_test3_labels():
    F, l = test3_problem()
    NP, NF = F.shape
    xstar = np.where(np.arange(NF) < NF // 2, 1.0, -1.0).astype(np.float32)[None, :]
    wb = F.view(np.uint8).reshape(NP, -1)
    y = orc.mul_mat_2d(orc.F32, wb, NP, NF, xstar, nth=8)[0]          # ggml_mul_mat(F, x): one value per F row
    assert np.abs(y - l).max() < 2e-2
    np.testing.assert_allclose(y, F.astype(np.float64) @ xstar[0].astype(np.float64), rtol=0, atol=1e-6)
    # thread count must not change a single bit (rows are independent, Ggml.cs:6130-6137)
    assert np.array_equal(y, orc.mul_mat_2d(orc.F32, wb, NP, NF, xstar, nth=3)[0])
    # the minimiser of sum((F.x - l)^2)/NP + 1e-5 |x|^2, using the oracle for F^T.(.) via mul_mat on the transpose
    A = F.astype(np.float64)
    x = np.linalg.solve(A.T @ A / NP + 1e-5 * np.eye(NF), A.T @ l.astype(np.float64) / NP)
    assert np.abs(x[:NF // 2] - 1).max() < 1e-2 and np.abs(x[NF // 2:] + 1).max() < 1e-2     # Test3/Program.cs:82-88

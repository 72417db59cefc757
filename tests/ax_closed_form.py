"""A known answer for the local Poisson operator that owes nothing to any implementation (used by tests/test_oracle.py on
the CPU and tests/test_nomp_api_gpu.py on the device).

Element = the reference cube under the shear X = xi + a eta + b zeta, Y = eta + c zeta, Z = zeta (det J = 1), so
G = J^-1 J^-T is a constant, FULL symmetric tensor (all six geometric factors in play); u = X Y Z is harmonic.  For a
constant G and u of degree <= N - 1 per direction the GLL quadrature of (grad phi, G grad u) is exact, and the divergence
theorem in reference coordinates gives
    (A u)_ijk = sum over the faces through node ijk of  +- (face quadrature weight) * (G grad_xi u) . e_face,
i.e. zero at interior nodes and the outward flux at boundary nodes.  Sensitive to the orientation of D, the order of the
six factors and the folded weights."""
import numpy as np


def sheared_element(n, x, a=0.3, b=-0.2, c=0.45):
    """x: the n GLL nodes.  Returns (u, g, want, interior mask) flattened in the layouts of include/nompk.h:
    u, want [k][j][i] (i fastest), g [6][k][j][i] with the quadrature weights folded in."""
    N = n - 1
    PN = np.polynomial.legendre.legval(x, [0.0] * N + [1.0])
    wt = 2.0 / (N * (N + 1) * PN ** 2)                               # GLL weights
    assert abs(wt.sum() - 2.0) < 1e-12
    J = np.array([[1.0, a, b], [0.0, 1.0, c], [0.0, 0.0, 1.0]])
    Ji = np.linalg.inv(J)
    G = Ji @ Ji.T
    k, j, i = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    xi, eta, zeta = x[i], x[j], x[k]
    X, Y, Z = xi + a * eta + b * zeta, eta + c * zeta, zeta
    u = X * Y * Z
    grad_X = np.stack([Y * Z, X * Z, X * Y])                         # physical gradient of XYZ
    grad_xi = np.einsum("ab,bkji->akji", J.T, grad_X)                # chain rule: d/dxi_a = sum_b dX_b/dxi_a d/dX_b
    flux = np.einsum("ab,bkji->akji", G, grad_xi)                    # G grad_xi u
    W3 = wt[k] * wt[j] * wt[i]
    g = np.stack([G[0, 0] * W3, G[0, 1] * W3, G[0, 2] * W3, G[1, 1] * W3, G[1, 2] * W3, G[2, 2] * W3])
    want = np.zeros_like(u)
    for axis, idx, other in ((0, i, wt[j] * wt[k]), (1, j, wt[i] * wt[k]), (2, k, wt[i] * wt[j])):
        want += np.where(idx == n - 1, other * flux[axis], 0.0) - np.where(idx == 0, other * flux[axis], 0.0)
    interior = (i > 0) & (i < n - 1) & (j > 0) & (j < n - 1) & (k > 0) & (k < n - 1)
    return (np.ascontiguousarray(u.ravel()), np.ascontiguousarray(g.ravel()), np.ascontiguousarray(want.ravel()),
            np.ascontiguousarray(interior.ravel()))

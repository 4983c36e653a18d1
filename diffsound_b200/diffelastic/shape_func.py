"""Lagrange shape functions on tetrahedra in barycentric coordinates.

API mirror of src/diffelastic/shape_func.py:3-108 (`get_shape_function`,
`get_shape_function_grad`) for the element orders the reference can actually
build (1 and 2; `TetMesh.to_high_order(3)` is unreachable there).

A quadratic node is described by the pair (l, m) of barycentric indices it
sits between: l == m is a corner with N = L_l (2 L_l - 1), l != m is the
mid-edge node with N = 4 L_l L_m.  The order of P2_NODES is the reference's
local node order [v0, m01, v1, m12, v2, m02, m03, m13, m23, v3].
"""
import torch

P2_NODES = ((0, 0), (0, 1), (1, 1), (1, 2), (2, 2), (2, 0), (0, 3), (1, 3), (2, 3), (3, 3))
NODES_PER_TET = {1: 4, 2: 10}
# local index of the four corner nodes (the only ones geometry depends on)
CORNER_LOCAL = {1: (0, 1, 2, 3), 2: (0, 2, 4, 9)}


def _check(order):
    if order not in NODES_PER_TET:
        raise NotImplementedError(f"element order {order} is not supported (1 or 2)")


def get_shape_function(L, order=1):
    """N_a at the points L (n, 4) -> (n, 4) or (n, 10)."""
    _check(order)
    if order == 1:
        return L
    cols = []
    for l, m in P2_NODES:
        if l == m:
            cols.append(L[:, l] * (2 * L[:, l] - 1))
        else:
            cols.append(4 * L[:, l] * L[:, m])
    return torch.stack(cols, dim=1)


def get_shape_function_grad(L, order=1):
    """dN_a/dL_l at the points L (n, 4) -> (n, nodes, 4)."""
    _check(order)
    n = L.shape[0]
    out = torch.zeros(n, NODES_PER_TET[order], 4, dtype=L.dtype, device=L.device)
    if order == 1:
        for a in range(4):
            out[:, a, a] = 1
        return out
    for a, (l, m) in enumerate(P2_NODES):
        if l == m:
            out[:, a, l] = 4 * L[:, l] - torch.ones_like(L[:, l])
        else:
            out[:, a, l] = 4 * L[:, m]
            out[:, a, m] = 4 * L[:, l]
    return out

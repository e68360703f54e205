"""Online diagnostics, API of core/online_diag.py:4-35.

VFwork: the point-wise work of the vortex force, u . (omega x U), and its sum over the interior.  It is
evaluated once per history snapshot (core/nyles.py:167-171), not on the hot path: the vortex force comes from
the same kernel the right-hand side uses (ny_vortex_force), the inner product is three elementwise torch
expressions in the operation order of the reference.
"""
from . import vortex_force as vortf

_AXIS = {"i": 2, "j": 1, "k": 0}          # position of a direction in the canonical (k, j, i) layout


class VFwork(object):
    """diagnose the point wise vortex-force work"""

    def __init__(self, model, grid):
        self.model = model
        self.grid = grid
        self.worksum = 0.0

    def compute(self):
        state = self.model.state
        dstate = self.model.timescheme.dstate
        self.compute_vortexforce(state, dstate)
        self.innerproduct(state.U, dstate.u, state.work)
        k0, k1, j0, j1, i0, i1 = self.grid.domainindices
        self.worksum = state.work.tensor[k0:k1, j0:j1, i0:i1].sum().item()

    def compute_vortexforce(self, state, dstate):
        for d in "ijk":
            dstate.u[d].tensor.zero_()
        vortf.vortex_force(state, dstate, self.model.orderVF)

    def innerproduct(self, vec1, vec2, result):
        """result = sum over the three directions of the face products averaged to the cell centre
        (online_diag.py:26-35): res[1:] += p[1:] + p[:-1] along each direction, p = u1*u2*0.5."""
        res = result.tensor
        res.zero_()
        for d in "ijk":
            a = _AXIS[d]
            n = res.shape[a]
            product = vec1[d].tensor * vec2[d].tensor * 0.5
            res.narrow(a, 1, n - 1).add_(product.narrow(a, 1, n - 1) + product.narrow(a, 0, n - 1))

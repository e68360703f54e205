"""Euler3d model, API of core/model_les_euler.py: the LES model without tracer advection, without
the buoyancy term and WITHOUT the halo fills of b and u before the projection
(diff against core/model_les.py at lines 36, 98-99, 131, 136)."""
from .model_les import LES as _LES


class LES(_LES):
    euler = True

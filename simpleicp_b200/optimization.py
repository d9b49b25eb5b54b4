"""Parameter containers returned by SimpleICP.run (API mirror of the reference's
python/simpleicp/optimization.py:291-382: ``Parameter`` and ``RigidBodyParameters``).

The optimisation itself (reference: SimpleICPOptimization, optimization.py:65-288) runs on the
GPU inside libsicp_b200 (csrc/reject_solve.cu); these dataclasses only carry its results.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List

import numpy as np

from . import mathutils

_RAD2DEG = 180 / np.pi
PARAMETER_NAMES = ("alpha1", "alpha2", "alpha3", "tx", "ty", "tz")


@dataclass
class Parameter:
    """One rigid-body parameter: initial / observed / estimated value, weight and sigma."""

    initial_value: float = np.nan
    observed_value: float = np.nan
    observation_weight: float = np.nan
    estimated_value: float = np.nan
    estimated_uncertainty: float = np.nan
    scale_for_logging: float = 1

    def _scaled(self, v):
        return v * self.scale_for_logging

    @property
    def initial_value_scaled(self):
        return self._scaled(self.initial_value)

    @property
    def observed_value_scaled(self):
        return self._scaled(self.observed_value)

    @property
    def estimated_value_scaled(self):
        return self._scaled(self.estimated_value)

    @property
    def estimated_uncertainty_scaled(self):
        return self._scaled(self.estimated_uncertainty)


def _angle():
    return Parameter(scale_for_logging=_RAD2DEG)


@dataclass
class RigidBodyParameters:
    """alpha1..3 (radians, logged in degrees) and tx, ty, tz."""

    alpha1: Parameter = field(default_factory=_angle)
    alpha2: Parameter = field(default_factory=_angle)
    alpha3: Parameter = field(default_factory=_angle)
    tx: Parameter = field(default_factory=Parameter)
    ty: Parameter = field(default_factory=Parameter)
    tz: Parameter = field(default_factory=Parameter)

    @property
    def H(self) -> np.ndarray:
        """Homogeneous transformation matrix of the estimated values."""
        v = self.get_parameter_attributes_as_list("estimated_value")
        return mathutils.create_homogeneous_transformation_matrix(
            mathutils.euler_angles_to_rotation_matrix(v[0], v[1], v[2]), v[3:6]
        )

    def set_parameter_attributes_from_list(self, attribute_name: str, array: List) -> None:
        for name, value in zip(PARAMETER_NAMES, array):
            setattr(getattr(self, name), attribute_name, value)

    def get_parameter_attributes_as_list(self, attribute_name: str) -> List:
        return [getattr(getattr(self, name), attribute_name) for name in PARAMETER_NAMES]

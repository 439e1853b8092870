"""Minimal ``gymnasium`` for machines where the real package is absent.

Put ``ia2c_b200/compat_gym`` on ``sys.path`` ONLY when gymnasium is not installed.  It provides exactly what the
reference's scripts touch — ``Env``, ``spaces.Discrete/Box``, ``envs.registration.register``, ``make_vec`` —
and ``make_vec`` of an id whose entry point is the drop-in ``Org`` returns the batched GPU environment
(``ia2c_b200.org_env.OrgVecEnv``): E envs stepped by one kernel launch with gymnasium 0.29.1's vector semantics
(TimeLimit truncation, same-step autoreset, float32 observations [E,6], float64 rewards [E]; ia2c.py:34-42,72,85).
When the real gymnasium IS installed, importing the drop-in ``Org`` module installs the same routing on
``gymnasium.make_vec`` (ia2c_b200.compat.Org).
"""
from . import spaces  # noqa: F401
from .core import Env, make, make_vec, register  # noqa: F401
from . import envs  # noqa: F401

__version__ = "0.29.1+ia2c_b200.shim"

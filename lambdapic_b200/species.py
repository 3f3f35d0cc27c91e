"""Species configuration: same names, fields and SI conversions as core/species.py:50-245 of the reference
(QED/spin variants are out of scope of the accelerated path)."""
from __future__ import annotations

import inspect
from dataclasses import dataclass, field
from typing import Callable, Literal

M_E = 9.1093837139e-31     # scipy.constants.m_e
M_P = 1.67262192595e-27    # scipy.constants.m_p
E_CHARGE = 1.602176634e-19  # scipy.constants.e


class EnableMixin:
    """core/utils/enable_mixin.py: disabled species are skipped by sort and push."""
    _enabled = True

    def enable(self):
        self._enabled = True

    def disable(self):
        self._enabled = False

    def is_enabled(self):
        return self._enabled


@dataclass(kw_only=True)
class Species(EnableMixin):
    name: str
    charge: int
    mass: float
    density: Callable | float | None = field(default=None)
    density_min: float = field(default=0)
    ppc: int | Callable = field(default=0)
    momentum: tuple | None = field(default=(None, None, None))
    polarization: tuple | None = field(default=None)
    pusher: Literal["boris", "photon", "boris+tbmt"] = field(default="boris")

    def __post_init__(self):
        if not isinstance(self.name, str) or not self.name:
            raise ValueError("species name must be a non-empty string")
        if self.pusher not in ("boris", "photon", "boris+tbmt"):
            raise ValueError(f"unknown pusher {self.pusher!r}")
        if self.pusher != "boris":
            raise NotImplementedError("only the Boris pusher is on the accelerated path (SURVEY.md 8)")
        if not callable(self.ppc) and self.ppc < 0:
            raise ValueError("ppc must be >= 0")
        self.m = self.mass * M_E
        self.q = self.charge * E_CHARGE
        self.density_jit = None
        self.ppc_jit = None
        self._aux_attrs: list[str] = []
        self._ispec: int | None = None

    @property
    def ispec(self) -> int:
        if self._ispec is None:
            raise ValueError("Species index is not set. Maybe not added via Simulation.add_species")
        return self._ispec

    @ispec.setter
    def ispec(self, value: int):
        self._ispec = value

    def is_compatible(self, dimension: int) -> bool:
        for func in (self.density, self.ppc):
            if inspect.isfunction(func) and func.__code__.co_argcount != dimension:
                return False
        return True

    @staticmethod
    def compile_profile(func_or_val, dimension: int):
        """Profile -> callable evaluated node by node on the host (the reference njit-compiles it,
        core/species.py:141-170); constants become constant functions."""
        if callable(func_or_val):
            if inspect.isfunction(func_or_val) and func_or_val.__code__.co_argcount != dimension:
                raise ValueError(f"function {func_or_val} must have {dimension} arguments")
            return func_or_val
        if isinstance(func_or_val, (int, float)):
            return lambda *xyz: func_or_val
        raise ValueError(f"Invalid profile {func_or_val}. Must be a function, int or float.")


@dataclass(kw_only=True)
class Electron(Species):
    name: str = field(default="electron", init=True)
    radiation: str | None = field(default=None, init=True)
    charge: int = field(default=-1, init=False)
    mass: float = field(default=1, init=False)

    def __post_init__(self):
        if self.radiation is not None:
            raise NotImplementedError("QED radiation is outside the accelerated path (SURVEY.md 2 #20)")
        super().__post_init__()


@dataclass(kw_only=True)
class Positron(Electron):
    name: str = field(default="positron", init=True)
    charge: int = field(default=1, init=False)


@dataclass(kw_only=True)
class Proton(Species):
    name: str = field(default="proton", init=True)
    charge: int = field(default=1, init=False)
    mass: float = field(default=M_P / M_E, init=False)

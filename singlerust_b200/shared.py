"""Public vocabulary of the reference API (src/shared/mod.rs:17-102), mirrored name for name."""
from __future__ import annotations

import enum
from dataclasses import dataclass


class Direction(enum.IntEnum):
    """src/shared/mod.rs:39-42 — identical discriminants (they cross the C ABI as int32)."""
    Row = 0
    Column = 1

    def is_row(self) -> bool:
        return self is Direction.Row


@dataclass(frozen=True)
class ComputationMode:
    """src/shared/mod.rs:25-28: ComputationMode::{Chunked(n), Whole}."""
    chunk: int | None = None

    @classmethod
    def Whole(cls):
        return cls(None)

    @classmethod
    def Chunked(cls, n: int):
        return cls(int(n))

    @property
    def is_whole(self):
        return self.chunk is None


@dataclass(frozen=True)
class FeatureSelection:
    """src/shared/mod.rs:17-23."""
    kind: str
    value: object = None

    @classmethod
    def HighlyVariableCol(cls, name: str):
        return cls("HighlyVariableCol", name)

    @classmethod
    def HighlyVariable(cls, n: int):
        return cls("HighlyVariable", int(n))

    @classmethod
    def Randomized(cls, n: int):
        return cls("Randomized", int(n))

    @classmethod
    def VarianceThreshold(cls, t: float):
        return cls("VarianceThreshold", float(t))

    @classmethod
    def None_(cls):
        return cls("None")


@dataclass(frozen=True)
class FlexValue:
    """src/shared/mod.rs:62-66 (used by the filters, which are outside the hot path; kept for API completeness)."""
    kind: str
    value: object = None

    @classmethod
    def Absolute(cls, v: int):
        return cls("Absolute", int(v))

    @classmethod
    def Relative(cls, v: float):
        return cls("Relative", float(v))

    @classmethod
    def None_(cls):
        return cls("None")

    def is_absolute(self):
        return self.kind == "Absolute"

    def is_relative(self):
        return self.kind == "Relative"

    def is_none(self):
        return self.kind == "None"

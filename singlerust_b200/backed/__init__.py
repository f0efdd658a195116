"""Mirror of the reference's src/backed front-end (out-of-core data)."""
from . import processing, statistics  # noqa: F401

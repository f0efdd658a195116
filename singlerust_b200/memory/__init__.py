"""Mirror of the reference's src/memory front-end over the B200 library."""
from . import processing, statistics  # noqa: F401

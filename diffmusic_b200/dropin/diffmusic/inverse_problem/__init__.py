"""Drop-in for the reference's `diffmusic.inverse_problem` (diffmusic/inverse_problem/__init__.py:1-11)."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)  # see ../schedulers/__init__.py

from diffmusic_b200.noise import GaussianNoise, PoissonNoise, get_noiser  # noqa: F401

from diffmusic_b200.noise import BaseNoise, GaussianNoise, PoissonNoise  # noqa: F401

"""``f_l = (-beta)^l / (2^l l!)`` -- same API as ``modulation_functions/diffusion_modulator.py:3-6``."""

import math


def diffusion_modulator(length, beta):
    return (-beta) ** length / (2 ** length * math.factorial(length))

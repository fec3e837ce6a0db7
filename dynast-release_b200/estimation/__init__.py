"""Operator layer mirroring dynast/estimation/__init__.py:2-10."""
from .alpha import estimate_alpha, read_alpha
from .p_c import estimate_p_c, read_p_c
from .p_e import estimate_p_e, estimate_p_e_control, estimate_p_e_nasc, read_p_e
from .pi import beta_mean, beta_mode, estimate_pi, guess_beta_parameters, read_pi

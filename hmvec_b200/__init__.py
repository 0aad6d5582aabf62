"""hmvec_b200 -- B200-native (sm_100a) implementation of the hmvec halo-model hot path.

Import surface mirrors the reference package (`hmvec/__init__.py:1` is `from .hmvec import *`): everything public in
`hmvec_b200.hmvec` is reachable as `hmvec_b200.<name>`.  Importing requires the built CUDA library
(hmvec_b200/libhmvec_b200.so); there is no CPU fallback.
"""
from .hmvec import *  # noqa: F401,F403
from .hmvec import HaloModel, DeviceCubes, duffy_concentration, R_from_M  # noqa: F401
from .cosmology import Cosmology, limber_integral, simpson_weights, Wkr, Wkr_taylor, a2z, get_eds_model  # noqa: F401
from .params import default_params, battaglia_defaults  # noqa: F401
from . import utils, zshard, tinker, fft, hostfuncs  # noqa: F401
from .fft import generic_profile_fft, fft_integral, analytic_fft_integral, uk_fft, uk_brute_force  # noqa: F401

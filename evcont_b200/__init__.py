"""evcont_b200: a B200-native engine behind evcont's FCI continuation API.

Module names mirror the reference package (``evcont/``): ``FCI_EVCont``,
``ab_initio_eigenvector_continuation``, ``ab_initio_gradients_loewdin``,
``electron_integral_utils``, ``MD_utils``.  Everything numerical runs in
``libevcont_b200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/evcont_b200.h``); there is no CPU fallback.
"""
__version__ = "0.1.0"

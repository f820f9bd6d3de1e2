"""rsa_b200 -- B200-native (sm_100a) rectified block-sparse attention: ctypes host layer over librsa_b200.so."""
from . import geometry, native, ops  # noqa: F401
from .native import RsaError  # noqa: F401

"""tloam_b200 -- B200-native (sm_100a) TLS scan-to-map registration behind T-LOAM's RegistrationInterface.

The product is the C-ABI CUDA library `libtloam_b200.so` (sources in tloam_b200/csrc, boundary in
include/tloam_b200.h).  This package is only the Python host-side mirror of the reference interface plus the
synthetic-scene generator used by tests and bench.py.  Importing it never touches oracle/.
"""
from .registration import BatchRegistration, Frame, LocalRegistration, RegistrationError, default_config  # noqa: F401

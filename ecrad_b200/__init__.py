"""ecrad_b200: B200-native (sm_100a CUDA) implementation of ecRad's per-column radiative-transfer hot path behind the
reference's radiation_interface API.  The compute lives in libecrad_b200.so (ecrad_b200/csrc); this package is the
host-side mirror of the reference interface (config, inputs, setup_radiation/radiation)."""

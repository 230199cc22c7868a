"""dcfa_b200: host side of the B200-native DCFA-YOLO inference hot path (plan compiler, weight packer,
ctypes binding of lib/libdcfa_b200.so).  The drop-in modules live next to this package in `nets/` and `utils/`."""
__version__ = "0.1.0"

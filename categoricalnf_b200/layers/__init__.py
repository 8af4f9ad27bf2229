"""Drop-in modules mirroring ``layers/`` of the reference (same class names, constructor
signatures, parameter/buffer names and return conventions), backed by the CUDA kernels."""

"""Drop-in replacement for the reference's ``pyfcd`` package (same module, class and method
names: pyfcd/fcd.py, pyfcd/fourier.py, pyfcd/carriers.py) running on fcd_b200's CUDA kernels."""

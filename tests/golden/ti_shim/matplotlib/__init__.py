"""Empty stand-in: the reference's sph_base.py imports `matplotlib.pyplot.axis` and never uses it."""

"""Import stub: usps_mnist.py:10 imports matplotlib.pyplot and never uses it; it is not installed here."""

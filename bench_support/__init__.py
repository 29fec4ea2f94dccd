"""Synthetic populations and data sets for the benchmark and the tests (NumPy restatements of the
reference's samplers).  Test / measurement support, not part of the product package."""

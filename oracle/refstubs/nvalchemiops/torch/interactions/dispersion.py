dftd3 = None  # never called with deterministic=True (aimnet/calculators/calculator.py:1002-1008)

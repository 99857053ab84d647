# never called with deterministic=True and method in {simple, dsf} (aimnet/calculators/calculator.py:971-975)
dsf_coulomb = ewald_summation = particle_mesh_ewald = None
estimate_ewald_parameters = estimate_pme_parameters = None

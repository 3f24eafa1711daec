/* Build shim: see gsl_rng.h in this directory. */

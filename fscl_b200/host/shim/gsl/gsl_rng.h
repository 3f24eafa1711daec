/* Build shim.  The reference's fscl.c includes <gsl/gsl_rng.h> and allocates one generator that is
 * never read (fscl.c:136-137); GSL is not in this image.  These no-op declarations let the
 * unmodified fscl.c compile; nothing here performs arithmetic. */
#ifndef FSCL_B200_GSL_RNG_SHIM_H
#define FSCL_B200_GSL_RNG_SHIM_H
typedef struct { int unused; } gsl_rng;
typedef struct { int unused; } gsl_rng_type;
static const gsl_rng_type *const gsl_rng_taus = 0;
static inline gsl_rng *gsl_rng_alloc(const gsl_rng_type *t) { (void)t; return (gsl_rng *)0; }
static inline void gsl_rng_set(gsl_rng *r, unsigned long s) { (void)r; (void)s; }
#endif

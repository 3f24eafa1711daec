/* Header shim (test infrastructure, not product code).
 * The reference's fscl path includes GSL headers but never draws from GSL
 * (/root/reference/fscl.c:136-137 allocates an RNG that is never read;
 * /root/reference/scan-chromosome.c:352 is a comment).  GSL is absent from
 * this image, so these no-op declarations let the unmodified reference
 * sources compile.  Nothing here performs arithmetic. */
#ifndef FSCL_ORACLE_GSL_RNG_SHIM_H
#define FSCL_ORACLE_GSL_RNG_SHIM_H
typedef struct { int unused; } gsl_rng;
typedef struct { int unused; } gsl_rng_type;
static const gsl_rng_type *const gsl_rng_taus = 0;
static inline gsl_rng *gsl_rng_alloc(const gsl_rng_type *t) { (void)t; return (gsl_rng *)0; }
static inline void gsl_rng_set(gsl_rng *r, unsigned long s) { (void)r; (void)s; }
#endif

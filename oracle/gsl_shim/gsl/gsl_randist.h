/* ORACLE / TEST INFRASTRUCTURE ONLY. See gsl_cdf.h in this directory. */
#ifndef MM_ORACLE_GSL_RANDIST_H
#define MM_ORACLE_GSL_RANDIST_H
#include "gsl_cdf.h"
#endif

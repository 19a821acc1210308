/*
 * usac_oracle_essential.cpp - five-point essential-matrix solver of the CPU oracle (placeholder until the
 * essential row of SURVEY.md section 8a is built). TEST INFRASTRUCTURE ONLY (see usac_oracle.h).
 */
#include "oracle_internal.h"

int orc_solve_essential5(const float*, const int*, float*) { return 0; }

/* randfill.h -- see randfill.cpp */
#pragma once
#include <stdint.h>
/* out[i] = (float)rand() / RAND_MAX for i < n, leaving the libc generator exactly where n calls of rand() would.
 * Returns true when the fast (lock-free, table-level) path was used. */
bool rand_fill(float *out, uint64_t n);

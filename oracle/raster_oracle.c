/*
 * raster_oracle.c -- CPU restatement of the reference rasterizer (fp32 + fp64 builds).
 * TEST INFRASTRUCTURE ONLY: linked into oracle/liboracle.so, which only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load.
 * Exports orc_*_32 (float, reference operation order) and orc_*_64 (double).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define REAL float
#define SUF 32
#define IS_FLOAT 1
#include "raster_oracle_impl.h"
#undef REAL
#undef SUF
#undef IS_FLOAT

#define REAL double
#define SUF 64
#define IS_FLOAT 0
#include "raster_oracle_impl.h"
#undef REAL
#undef SUF
#undef IS_FLOAT

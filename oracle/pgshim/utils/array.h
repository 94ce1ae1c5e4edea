/* pgshim/utils/array.h -- TEST INFRASTRUCTURE ONLY: the few array calls the GPU glue makes.  An "array" is a flat vector of
 * datums; deconstruct_array() keeps PostgreSQL's signature. */
#ifndef PGSHIM_ARRAY_H
#define PGSHIM_ARRAY_H
#include "postgres.h"
typedef struct ArrayType { int nelems; unsigned int elemtype; Datum elems[]; } ArrayType;
#define PG_GETARG_ARRAYTYPE_P(n) ((ArrayType *) PG_GETARG_POINTER(n))
#define ARR_ELEMTYPE(a) ((a)->elemtype)
static inline void deconstruct_array(ArrayType *a, unsigned int elmtype, int elmlen, bool elmbyval, char elmalign, Datum **elemsp,
									 bool **nullsp, int *nelemsp)
{
	(void) elmtype; (void) elmlen; (void) elmbyval; (void) elmalign;
	*elemsp = a->elems;
	if (nullsp) *nullsp = NULL;
	*nelemsp = a->nelems;
}
#endif

/* pgshim/funcapi.h -- TEST INFRASTRUCTURE ONLY: value-per-call SRF protocol stand-in. */
#ifndef PGSHIM_FUNCAPI_H
#define PGSHIM_FUNCAPI_H
#include "fmgr.h"

typedef struct FuncCallContext
{
	uint64_t call_cntr;
	uint64_t max_calls;
	void *user_fctx;
	MemoryContext multi_call_memory_ctx;
} FuncCallContext;

extern FuncCallContext *pgshim_srf_firstcall_init(FunctionCallInfo fcinfo);

#define SRF_IS_FIRSTCALL() (fcinfo->srf_ctx == NULL)
#define SRF_FIRSTCALL_INIT() pgshim_srf_firstcall_init(fcinfo)
#define SRF_PERCALL_SETUP() (fcinfo->srf_ctx)
#define SRF_RETURN_NEXT(funcctx, result) \
	do { (funcctx)->call_cntr++; fcinfo->srf_done = false; return (result); } while (0)
#define SRF_RETURN_DONE(funcctx) \
	do { fcinfo->srf_done = true; fcinfo->isnull = true; return (Datum) 0; } while (0)

/* ---- composite results (used by the GPU glue, kmer-extension_b200/pgglue/kmer_gpu.c; the reference itself never builds
 *      tuples).  A "tuple" here is just the array of its attribute datums. ---- */
typedef struct PgShimTupleDesc { int natts; } *TupleDesc;
typedef struct PgShimHeapTuple { int natts; Datum values[8]; bool nulls[8]; } *HeapTuple;
typedef enum TypeFuncClass { TYPEFUNC_SCALAR, TYPEFUNC_COMPOSITE, TYPEFUNC_OTHER } TypeFuncClass;
typedef unsigned int Oid;
extern TypeFuncClass pgshim_get_call_result_type(FunctionCallInfo fcinfo, Oid *resultTypeId, TupleDesc *resultTupleDesc);
extern HeapTuple pgshim_heap_form_tuple(TupleDesc desc, Datum *values, bool *isnull);
#define get_call_result_type(fcinfo, oidp, descp) pgshim_get_call_result_type(fcinfo, oidp, descp)
#define BlessTupleDesc(d) (d)
#define heap_form_tuple(desc, values, nulls) pgshim_heap_form_tuple(desc, values, nulls)
#define HeapTupleGetDatum(t) PointerGetDatum(t)

#endif

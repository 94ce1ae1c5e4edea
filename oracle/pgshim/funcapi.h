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

#endif

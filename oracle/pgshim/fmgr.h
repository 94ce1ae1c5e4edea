/* pgshim/fmgr.h -- TEST INFRASTRUCTURE ONLY: fmgr-V1 calling convention stand-in. */
#ifndef PGSHIM_FMGR_H
#define PGSHIM_FMGR_H
#include "postgres.h"

struct FuncCallContext;

typedef struct NullableDatum
{
	Datum value;
	bool isnull;
} NullableDatum;

typedef struct FunctionCallInfoBaseData
{
	struct FuncCallContext *srf_ctx; /* stands in for flinfo->fn_extra */
	bool srf_done;					 /* stands in for ReturnSetInfo.isDone == ExprEndResult */
	bool isnull;
	short nargs;
	NullableDatum args[4];
} FunctionCallInfoBaseData;
typedef FunctionCallInfoBaseData *FunctionCallInfo;

#define PG_FUNCTION_ARGS FunctionCallInfo fcinfo
#define PG_MODULE_MAGIC extern int pgshim_module_magic
#define PG_FUNCTION_INFO_V1(fn) extern Datum fn(PG_FUNCTION_ARGS)

#define PG_GETARG_DATUM(n) (fcinfo->args[n].value)
#define PG_GETARG_POINTER(n) DatumGetPointer(PG_GETARG_DATUM(n))
#define PG_GETARG_CSTRING(n) ((char *) PG_GETARG_POINTER(n))
#define PG_GETARG_VARLENA_P(n) ((struct varlena *) PG_GETARG_POINTER(n)) /* nothing is ever toasted here */
#define PG_GETARG_INT32(n) ((int32) PG_GETARG_DATUM(n))
#define PG_ARGISNULL(n) (fcinfo->args[n].isnull)

#define PG_RETURN_DATUM(x) return (x)
#define PG_RETURN_POINTER(x) return PointerGetDatum(x)
#define PG_RETURN_CSTRING(x) return PointerGetDatum(x)
#define PG_RETURN_INT32(x) return (Datum) (uint32) (int32) (x)
#define PG_RETURN_BOOL(x) return (Datum) ((x) ? 1 : 0)

#endif

/* pgshim: stand-in for PostgreSQL's parser/parse_func.h (test infrastructure only): function lookup in the driver's tiny catalog. */
#ifndef PGSHIM_PARSE_FUNC_H
#define PGSHIM_PARSE_FUNC_H
#include "nodes/pgshim_nodes.h"
extern Oid LookupFuncName(List *funcname, int nargs, const Oid *argtypes, bool missing_ok);
#endif

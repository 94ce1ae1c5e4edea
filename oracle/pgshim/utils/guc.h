/* pgshim: stand-in for PostgreSQL's utils/guc.h (test infrastructure only): a custom bool variable is just its C variable. */
#ifndef PGSHIM_GUC_H
#define PGSHIM_GUC_H
#include "postgres.h"
typedef enum GucContext { PGC_INTERNAL, PGC_POSTMASTER, PGC_SIGHUP, PGC_SU_BACKEND, PGC_BACKEND, PGC_SUSET, PGC_USERSET } GucContext;
typedef bool (*GucBoolCheckHook)(bool *newval, void **extra, int source);
typedef void (*GucBoolAssignHook)(bool newval, void *extra);
typedef const char *(*GucShowHook)(void);
extern void DefineCustomBoolVariable(const char *name, const char *short_desc, const char *long_desc, bool *valueAddr, bool bootValue,
									 GucContext context, int flags, GucBoolCheckHook check_hook, GucBoolAssignHook assign_hook, GucShowHook show_hook);
#endif

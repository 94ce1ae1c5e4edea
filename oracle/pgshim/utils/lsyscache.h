/* pgshim: stand-in for PostgreSQL's utils/lsyscache.h (test infrastructure only). */
#ifndef PGSHIM_LSYSCACHE_H
#define PGSHIM_LSYSCACHE_H
#include "nodes/pgshim_nodes.h"
extern Oid get_array_type(Oid typid);
#endif

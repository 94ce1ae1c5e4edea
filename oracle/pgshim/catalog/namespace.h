/* pgshim: stand-in for PostgreSQL's catalog/namespace.h (test infrastructure only): type lookup in the driver's tiny catalog. */
#ifndef PGSHIM_NAMESPACE_H
#define PGSHIM_NAMESPACE_H
#include "nodes/pgshim_nodes.h"
extern Oid TypenameGetTypid(const char *typname);
#endif

/* pgshim: stand-in for PostgreSQL's catalog/pg_type.h (test infrastructure only).  kmer.c includes it and uses nothing of it;
 * the planner hook of the GPU glue needs three built-in type OIDs (values as in pg_type.dat). */
#ifndef PGSHIM_PG_TYPE_H
#define PGSHIM_PG_TYPE_H
#define INT8OID 20
#define INT4OID 23
#define RECORDOID 2249
#endif

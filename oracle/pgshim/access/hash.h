/* pgshim/access/hash.h -- TEST INFRASTRUCTURE ONLY.
 * hash_any's value never reaches query results (it only orders HashAggregate buckets), so any
 * byte hash is a faithful stand-in; the driver uses it for its own hash-aggregate table. */
#ifndef PGSHIM_HASH_H
#define PGSHIM_HASH_H
#include "postgres.h"
extern Datum hash_any(const unsigned char *k, int keylen);
#endif

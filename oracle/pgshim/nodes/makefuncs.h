/* pgshim: stand-in for PostgreSQL's nodes/makefuncs.h (test infrastructure only): see pgshim_nodes.h */
#include "nodes/pgshim_nodes.h"

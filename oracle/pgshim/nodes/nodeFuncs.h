/* pgshim: stand-in for PostgreSQL's nodes/nodeFuncs.h (test infrastructure only): see pgshim_nodes.h */
#include "nodes/pgshim_nodes.h"

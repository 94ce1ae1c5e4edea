/* pgshim: stand-in for PostgreSQL's optimizer/planner.h (test infrastructure only): the planner hook chain.
 * standard_planner() here "plans" by handing back the Query it was given, so a test can look at what the hook made of it. */
#ifndef PGSHIM_PLANNER_H
#define PGSHIM_PLANNER_H
#include "nodes/pgshim_nodes.h"
typedef PlannedStmt *(*planner_hook_type)(Query *parse, const char *query_string, int cursorOptions, ParamListInfo boundParams);
extern planner_hook_type planner_hook;
extern PlannedStmt *standard_planner(Query *parse, const char *query_string, int cursorOptions, ParamListInfo boundParams);
#endif

/*
 * oracle/pgshim_nodes.c -- TEST INFRASTRUCTURE ONLY: runtime of pgshim/nodes/pgshim_nodes.h and of the hook-related stand-in
 * headers (planner hook chain, a tiny name -> OID catalog the test driver fills, custom GUC registration).  Linked into
 * tests/c/hook_driver only; palloc comes from oracle/_ref/libkmer_ref.so (ref_driver.c) like for every other shim user.
 */
#include "postgres.h"
#include "catalog/namespace.h"
#include "catalog/pg_type.h"
#include "nodes/pgshim_nodes.h"
#include "optimizer/planner.h"
#include "parser/parse_func.h"
#include "utils/guc.h"
#include "utils/lsyscache.h"

void *pgshim_new_node(Size size, NodeTag tag)
{
	Node *n = (Node *) palloc(size);
	memset(n, 0, size);
	n->type = tag;
	return n;
}

static List *list_grow(List *l)
{
	if (l == NIL)
	{
		l = (List *) pgshim_new_node(sizeof(List), T_List);
		l->max_length = 4;
		l->elements = (ListCell *) palloc(sizeof(ListCell) * 4);
	}
	else if (l->length == l->max_length)
	{
		ListCell *e = (ListCell *) palloc(sizeof(ListCell) * (size_t) l->max_length * 2);
		memcpy(e, l->elements, sizeof(ListCell) * (size_t) l->length);
		l->elements = e;
		l->max_length *= 2;
	}
	return l;
}

List *lappend(List *l, void *datum)
{
	l = list_grow(l);
	l->elements[l->length++].ptr_value = datum;
	return l;
}

List *lappend_oid(List *l, Oid datum)
{
	l = list_grow(l);
	l->elements[l->length++].oid_value = datum;
	return l;
}

String *makeString(char *str)
{
	String *s = makeNode(String);
	s->sval = str;
	return s;
}

Var *makeVar(int varno, AttrNumber varattno, Oid vartype, int32 vartypmod, Oid varcollid, Index varlevelsup)
{
	Var *v = makeNode(Var);
	v->varno = varno; v->varattno = varattno; v->vartype = vartype; v->vartypmod = vartypmod; v->varcollid = varcollid;
	v->varlevelsup = varlevelsup; v->location = -1;
	return v;
}

TargetEntry *makeTargetEntry(Expr *expr, AttrNumber resno, char *resname, bool resjunk)
{
	TargetEntry *t = makeNode(TargetEntry);
	t->expr = expr; t->resno = resno; t->resname = resname; t->resjunk = resjunk;
	return t;
}

FuncExpr *makeFuncExpr(Oid funcid, Oid rettype, List *args, Oid funccollid, Oid inputcollid, CoercionForm fformat)
{
	FuncExpr *f = makeNode(FuncExpr);
	f->funcid = funcid; f->funcresulttype = rettype; f->funcretset = false; f->funcvariadic = false; f->funcformat = fformat;
	f->funccollid = funccollid; f->inputcollid = inputcollid; f->args = args; f->location = -1;
	return f;
}

Alias *makeAlias(const char *aliasname, List *colnames)
{
	Alias *a = makeNode(Alias);
	a->aliasname = pstrdup(aliasname);
	a->colnames = colnames;
	return a;
}

FromExpr *makeFromExpr(List *fromlist, Node *quals)
{
	FromExpr *f = makeNode(FromExpr);
	f->fromlist = fromlist;
	f->quals = quals;
	return f;
}

Oid exprType(const Node *expr)
{
	if (!expr) return InvalidOid;
	switch (nodeTag(expr))
	{
	case T_Var: return ((const Var *) expr)->vartype;
	case T_Const: return ((const Const *) expr)->consttype;
	case T_FuncExpr: return ((const FuncExpr *) expr)->funcresulttype;
	case T_Aggref: return ((const Aggref *) expr)->aggtype;
	case T_ArrayExpr: return ((const ArrayExpr *) expr)->array_typeid;
	default: return InvalidOid;
	}
}

/* ---- the tiny catalog ---- */
typedef struct { const char *name; Oid oid, array_oid; } ShimType;
typedef struct { const char *name; int nargs; Oid args[4]; Oid oid; } ShimFunc;
static ShimType shim_types[16];
static ShimFunc shim_funcs[16];
static int n_types, n_funcs;

void pgshim_catalog_reset(void) { n_types = n_funcs = 0; }
void pgshim_catalog_add_type(const char *name, Oid oid, Oid array_oid)
{
	shim_types[n_types].name = name; shim_types[n_types].oid = oid; shim_types[n_types].array_oid = array_oid; n_types++;
}
void pgshim_catalog_add_func(const char *name, int nargs, const Oid *args, Oid oid)
{
	shim_funcs[n_funcs].name = name; shim_funcs[n_funcs].nargs = nargs; shim_funcs[n_funcs].oid = oid;
	for (int i = 0; i < nargs; i++) shim_funcs[n_funcs].args[i] = args[i];
	n_funcs++;
}

Oid TypenameGetTypid(const char *typname)
{
	for (int i = 0; i < n_types; i++)
		if (!strcmp(shim_types[i].name, typname)) return shim_types[i].oid;
	return InvalidOid;
}

Oid get_array_type(Oid typid)
{
	for (int i = 0; i < n_types; i++)
		if (shim_types[i].oid == typid) return shim_types[i].array_oid;
	return InvalidOid;
}

Oid LookupFuncName(List *funcname, int nargs, const Oid *argtypes, bool missing_ok)
{
	const char *name = strVal(list_nth(funcname, list_length(funcname) - 1));
	for (int i = 0; i < n_funcs; i++)
	{
		if (strcmp(shim_funcs[i].name, name) || shim_funcs[i].nargs != nargs) continue;
		int same = 1;
		for (int a = 0; a < nargs; a++) same &= shim_funcs[i].args[a] == argtypes[a];
		if (same) return shim_funcs[i].oid;
	}
	if (!missing_ok)
		ereport(ERROR, (errcode(ERRCODE_INTERNAL_ERROR), errmsg("function %s does not exist", name)));
	return InvalidOid;
}

/* ---- planner hook chain, GUC ---- */
planner_hook_type planner_hook = NULL;

PlannedStmt *standard_planner(Query *parse, const char *query_string, int cursorOptions, ParamListInfo boundParams)
{
	PlannedStmt *p = makeNode(PlannedStmt);
	(void) query_string; (void) cursorOptions; (void) boundParams;
	p->pgshim_query = parse;
	return p;
}

static struct { const char *name; bool *addr; } shim_gucs[8];
static int n_gucs;

void DefineCustomBoolVariable(const char *name, const char *short_desc, const char *long_desc, bool *valueAddr, bool bootValue, GucContext context,
							  int flags, GucBoolCheckHook check_hook, GucBoolAssignHook assign_hook, GucShowHook show_hook)
{
	(void) short_desc; (void) long_desc; (void) context; (void) flags; (void) check_hook; (void) assign_hook; (void) show_hook;
	*valueAddr = bootValue;
	shim_gucs[n_gucs].name = name;
	shim_gucs[n_gucs].addr = valueAddr;
	n_gucs++;
}

/* SET name = on|off */
bool pgshim_set_bool_guc(const char *name, bool value)
{
	for (int i = 0; i < n_gucs; i++)
		if (!strcmp(shim_gucs[i].name, name)) { *shim_gucs[i].addr = value; return true; }
	return false;
}

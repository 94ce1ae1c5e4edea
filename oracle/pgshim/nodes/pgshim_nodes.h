/*
 * pgshim/nodes/pgshim_nodes.h -- TEST INFRASTRUCTURE ONLY.
 *
 * The part of PostgreSQL's parse-tree node system that the planner hook of the GPU glue
 * (kmer-extension_b200/pgglue/kmer_gpu_hook.c) touches, so that the hook compiles here (no PostgreSQL in the image)
 * and RUNS against hand-built analyzed Query trees (tests/c/hook_driver.c).  Struct and field names are PostgreSQL's
 * (nodes/parsenodes.h, nodes/primnodes.h, nodes/pg_list.h of PostgreSQL 13-16); only the fields the hook or the driver reads or
 * writes exist.  Nothing of PostgreSQL's source is copied: these are declarations restated from its public extension API.
 */
#ifndef PGSHIM_NODES_H
#define PGSHIM_NODES_H
#include "postgres.h"

typedef unsigned int Oid;
#define InvalidOid ((Oid) 0)
typedef unsigned int Index;
typedef int16_t AttrNumber;
typedef uint32_t bits32;

typedef enum NodeTag
{
	T_Invalid = 0, T_List, T_String, T_Alias, T_Var, T_Const, T_FuncExpr, T_Aggref, T_SubLink, T_ArrayExpr, T_TargetEntry,
	T_RangeTblRef, T_FromExpr, T_Query, T_RangeTblEntry, T_RangeTblFunction, T_SortGroupClause, T_OpExpr, T_PlannedStmt
} NodeTag;

typedef struct Node { NodeTag type; } Node;
#define nodeTag(n) (((const Node *) (n))->type)
#define IsA(n, t) ((n) != NULL && nodeTag(n) == T_##t)
extern void *pgshim_new_node(Size size, NodeTag tag);
#define makeNode(t) ((t *) pgshim_new_node(sizeof(t), T_##t))
#define castNode(t, n) ((t *) (n))

/* ---- pg_list.h ---- */
typedef union ListCell { void *ptr_value; int int_value; Oid oid_value; } ListCell;
typedef struct List { NodeTag type; int length; int max_length; ListCell *elements; } List;
#define NIL ((List *) NULL)
static inline int list_length(const List *l) { return l ? l->length : 0; }
extern List *lappend(List *list, void *datum);
extern List *lappend_oid(List *list, Oid datum);
#define list_make1(a) lappend(NIL, (a))
#define list_make2(a, b) lappend(lappend(NIL, (a)), (b))
#define lfirst(lc) ((lc)->ptr_value)
#define lfirst_oid(lc) ((lc)->oid_value)
#define lfirst_node(t, lc) ((t *) lfirst(lc))
#define list_nth(l, n) ((l)->elements[n].ptr_value)
#define linitial(l) list_nth(l, 0)
#define lsecond(l) list_nth(l, 1)
#define linitial_node(t, l) ((t *) linitial(l))
#define foreach(cell, lst) \
	for (int cell##__i = 0; ((cell) = (cell##__i < list_length(lst)) ? &(lst)->elements[cell##__i] : NULL) != NULL; cell##__i++)

/* ---- value.h / primnodes.h ---- */
typedef struct String { NodeTag type; char *sval; } String;
extern String *makeString(char *str);
#define strVal(v) (((String *) (v))->sval)

typedef struct Alias { NodeTag type; char *aliasname; List *colnames; } Alias;
typedef struct Expr { NodeTag type; } Expr;

typedef struct Var
{
	Expr xpr;
	int varno;
	AttrNumber varattno;
	Oid vartype;
	int32 vartypmod;
	Oid varcollid;
	Index varlevelsup;
	int location;
} Var;

typedef struct Const
{
	Expr xpr;
	Oid consttype;
	int32 consttypmod;
	Oid constcollid;
	int constlen;
	Datum constvalue;
	bool constisnull;
	bool constbyval;
	int location;
} Const;

typedef enum CoercionForm { COERCE_EXPLICIT_CALL, COERCE_EXPLICIT_CAST, COERCE_IMPLICIT_CAST } CoercionForm;

typedef struct FuncExpr
{
	Expr xpr;
	Oid funcid;
	Oid funcresulttype;
	bool funcretset;
	bool funcvariadic;
	CoercionForm funcformat;
	Oid funccollid;
	Oid inputcollid;
	List *args;
	int location;
} FuncExpr;

typedef struct Aggref
{
	Expr xpr;
	Oid aggfnoid;
	Oid aggtype;
	List *aggdirectargs;
	List *args;
	List *aggorder;
	List *aggdistinct;
	Expr *aggfilter;
	bool aggstar;
	bool aggvariadic;
	char aggkind;
	Index agglevelsup;
	int location;
} Aggref;

typedef enum SubLinkType { EXISTS_SUBLINK, ALL_SUBLINK, ANY_SUBLINK, ROWCOMPARE_SUBLINK, EXPR_SUBLINK, MULTIEXPR_SUBLINK, ARRAY_SUBLINK, CTE_SUBLINK } SubLinkType;
typedef struct SubLink
{
	Expr xpr;
	SubLinkType subLinkType;
	int subLinkId;
	Node *testexpr;
	List *operName;
	Node *subselect;
	int location;
} SubLink;

typedef struct ArrayExpr
{
	Expr xpr;
	Oid array_typeid;
	Oid array_collid;
	Oid element_typeid;
	List *elements;
	bool multidims;
	int location;
} ArrayExpr;

typedef struct OpExpr { Expr xpr; Oid opno; List *args; } OpExpr; /* only so that the driver can build a WHERE clause */

typedef struct TargetEntry
{
	Expr xpr;
	Expr *expr;
	AttrNumber resno;
	char *resname;
	Index ressortgroupref;
	Oid resorigtbl;
	AttrNumber resorigcol;
	bool resjunk;
} TargetEntry;

typedef struct RangeTblRef { NodeTag type; int rtindex; } RangeTblRef;
typedef struct FromExpr { NodeTag type; List *fromlist; Node *quals; } FromExpr;

/* ---- parsenodes.h ---- */
typedef enum CmdType { CMD_UNKNOWN, CMD_SELECT, CMD_UPDATE, CMD_INSERT, CMD_DELETE, CMD_UTILITY } CmdType;
typedef enum QuerySource { QSRC_ORIGINAL, QSRC_PARSER, QSRC_INSTEAD_RULE } QuerySource;
typedef enum RTEKind { RTE_RELATION, RTE_SUBQUERY, RTE_JOIN, RTE_FUNCTION, RTE_TABLEFUNC, RTE_VALUES, RTE_CTE } RTEKind;

typedef struct Query
{
	NodeTag type;
	CmdType commandType;
	QuerySource querySource;
	bool canSetTag;
	Node *utilityStmt;
	int resultRelation;
	bool hasAggs;
	bool hasWindowFuncs;
	bool hasTargetSRFs;
	bool hasSubLinks;
	bool hasDistinctOn;
	bool hasRecursive;
	bool hasModifyingCTE;
	bool hasForUpdate;
	bool hasRowSecurity;
	List *cteList;
	List *rtable;
	List *rteperminfos; /* PostgreSQL 16+ */
	FromExpr *jointree;
	List *targetList;
	List *returningList;
	List *groupClause;
	List *groupingSets;
	Node *havingQual;
	List *windowClause;
	List *distinctClause;
	List *sortClause;
	Node *limitOffset;
	Node *limitCount;
	List *rowMarks;
	Node *setOperations;
} Query;

typedef struct RangeTblEntry
{
	NodeTag type;
	RTEKind rtekind;
	Oid relid;
	char relkind;
	Index perminfoindex; /* PostgreSQL 16+ */
	Query *subquery;
	List *functions;
	bool funcordinality;
	Alias *alias;
	Alias *eref;
	bool lateral;
	bool inh;
	bool inFromCl;
} RangeTblEntry;

typedef struct RangeTblFunction
{
	NodeTag type;
	Node *funcexpr;
	int funccolcount;
	List *funccolnames;
	List *funccoltypes;
	List *funccoltypmods;
	List *funccolcollations;
} RangeTblFunction;

typedef struct SortGroupClause
{
	NodeTag type;
	Index tleSortGroupRef;
	Oid eqop;
	Oid sortop;
	bool nulls_first;
	bool hashable;
} SortGroupClause;

typedef struct PlannedStmt { NodeTag type; Query *pgshim_query; /* the shim's "plan" is the Query it was handed */ } PlannedStmt;
typedef struct ParamListInfoData *ParamListInfo;

/* ---- makefuncs.h ---- */
extern Var *makeVar(int varno, AttrNumber varattno, Oid vartype, int32 vartypmod, Oid varcollid, Index varlevelsup);
extern TargetEntry *makeTargetEntry(Expr *expr, AttrNumber resno, char *resname, bool resjunk);
extern FuncExpr *makeFuncExpr(Oid funcid, Oid rettype, List *args, Oid funccollid, Oid inputcollid, CoercionForm fformat);
extern Alias *makeAlias(const char *aliasname, List *colnames);
extern FromExpr *makeFromExpr(List *fromlist, Node *quals);
extern Oid exprType(const Node *expr);

#define PG_VERSION_NUM 160000

#endif
